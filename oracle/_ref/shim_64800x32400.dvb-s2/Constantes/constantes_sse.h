#include "/root/reference/code/x86/Constantes/64800x32400.dvb-s2/constantes_sse.h"
