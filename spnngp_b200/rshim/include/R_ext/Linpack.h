#ifndef RSHIM_LINPACK_H
#define RSHIM_LINPACK_H
#include "RS.h"
#endif
