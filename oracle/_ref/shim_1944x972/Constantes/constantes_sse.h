#include "/root/reference/code/x86/Constantes/1944x972/constantes_sse.h"
