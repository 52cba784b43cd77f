"""CPU suite, part 3: the multi-GPU host logic (frame sharding, counter reduction) over a world-size-2 gloo group.
The decode itself is replaced by the oracle here (there is no GPU and the product has no CPU path) — what is under test is that
shards tile the batch exactly and that the reduced BER/FER counters equal the single-process ones."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from ldpcgputegra_b200.sharding import shard
from _helpers import ROOT


def test_shard_partition_properties():
    for total in (0, 1, 2, 7, 65536, 65537, 100003):
        for world in (1, 2, 3, 4, 8):
            parts = [shard(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        shard(10, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    import numpy as np, torch, torch.distributed as dist
    from ldpcgputegra_b200 import Code, default_params
    from ldpcgputegra_b200.sharding import shard, reduce_counters, max_over_ranks
    from _helpers import awgn_llr, oracle_decode
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    code = Code.load("576x288")
    total = 301
    llr = awgn_llr(code, total, 1.5, seed=9)              # every rank regenerates the same batch, then takes its slice
    f0, cnt = shard(total, world, rank)
    hard = oracle_decode(code, default_params(), llr[f0:f0 + cnt], 5, want_state=False)["hard"]
    info = hard[:, :code.k_info]
    red = reduce_counters([cnt, int(info.sum()), int(info.any(axis=1).sum())], dist)
    tmax = max_over_ranks(1.0 + rank, dist)
    if rank == 0:
        print("RESULT " + json.dumps(dict(red=red, tmax=tmax)))
    dist.destroy_process_group()
""")


def test_two_rank_gloo_counters_match_single_process(tmp_path):
    import json
    from ldpcgputegra_b200 import Code, default_params
    from _helpers import awgn_llr, oracle_decode
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    got = json.loads(line[7:])
    code = Code.load("576x288")
    llr = awgn_llr(code, 301, 1.5, seed=9)
    info = oracle_decode(code, default_params(), llr, 5, want_state=False)["hard"][:, :code.k_info]
    assert got["red"] == [301, int(info.sum()), int(info.any(axis=1).sum())]
    assert got["tmax"] == 2.0
