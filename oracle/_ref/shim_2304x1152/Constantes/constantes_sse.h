#include "/root/reference/code/x86/Constantes/2304x1152/constantes_sse.h"
