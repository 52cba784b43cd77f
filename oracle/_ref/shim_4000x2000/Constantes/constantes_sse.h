#include "/root/reference/code/x86/Constantes/4000x2000/constantes_sse.h"
