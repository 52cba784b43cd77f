#include "/root/reference/code/x86/Constantes/576x288/constantes_sse.h"
