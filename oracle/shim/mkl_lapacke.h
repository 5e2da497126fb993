/* TEST INFRASTRUCTURE — see mkl.h in this directory (the reference includes both names). */
#include "mkl.h"
