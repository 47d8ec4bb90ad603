/* intp.c */
#include "mus_oracle.h"
