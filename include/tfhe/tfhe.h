/* <tfhe/tfhe.h> — umbrella header of the TFHE-compatible front end of peba1-b200 (see tfhe_core.h). */
#ifndef PEBA1_TFHE_H
#define PEBA1_TFHE_H
#include "tfhe_core.h"
#include "tfhe_gate_bootstrapping_functions.h"
#include "../peba1_host.h"
#endif
