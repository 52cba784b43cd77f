#include "/root/reference/code/x86/Constantes/2048x384/constantes_sse.h"
