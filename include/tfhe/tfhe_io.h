/*
 * <tfhe/tfhe_io.h> — the reference includes this header (include/Math.h:5, include/Client.h:5) but calls none of its
 * functions, so serialisation is out of scope; the header exists so the reference compiles
 * unchanged.
 */
#ifndef PEBA1_TFHE_IO_H
#define PEBA1_TFHE_IO_H
#include "tfhe_core.h"
#endif
