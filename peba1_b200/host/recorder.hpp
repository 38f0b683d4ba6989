// recorder.hpp — the gate recorder / level scheduler between the TFHE-compatible API and libpeba1_cuda.
//
// boots* calls do not compute: they append SSA nodes.  A *wire* is one ciphertext VALUE (not a storage location):
// the reference reuses temporaries freely (src/Math.cpp:34-42 writes tmp[0] four times), aliases outputs with inputs
// (:272) and frees temporaries before their consumers have run (:47-49); with values named by wires none of that
// matters.  NOT / COPY / CONSTANT never launch anything: a handle is (wire << 1 | negated) and wire 1 is the constant
// false, so they are handle arithmetic.  Each node gets level = 1 + max(level of its operands); flush() executes
// level by level, one peba1_cuda_gates() call (blind-rotation launch + keyswitch launch) per level, assigning device
// pool slots to wires only while their values are live.
#pragma once

#include <stdint.h>

#include <vector>

#include "../../include/peba1_cuda.h"
#include "../../include/peba1_host.h"

namespace peba1 {

using Handle = uint64_t;  // 0: host value (only inside LweSample); else (engine index + 1) << 40 | (wire << 1) | neg
constexpr Handle kFalse = 2, kTrue = 3;
inline bool is_const(Handle h) { return (h >> 1) == 1; }
inline Handle negate(Handle h) { return h ^ 1; }

struct Engine;  // one per key material: CUDA context + recorder state

// engine of a cloud key (created on first use; aborts when no device / library error: there is no CPU fallback)
Engine* engine_of(const TFheGateBootstrappingCloudKeySet* bk);

// reference counting of wires held outside the recorder (samples, circuit-builder temporaries)
void retain(Handle h);
void release(Handle h);

// recording
Handle record_gate(Engine* e, int op, Handle a, Handle b);           // 2-input bootstrapped gate
Handle record_mux(Engine* e, Handle sel, Handle if1, Handle if0);    // 2 bootstraps + 1 keyswitch
Handle record_gate3(Engine* e, int op, Handle a, Handle b, Handle c);  // XOR3 / MAJ3: 1 bootstrap + 1 keyswitch
Handle wire_from_host(Engine* e, const int32_t* words);              // queues an upload; returns a handle with 1 ref
void flush(Engine* e);
void fetch(Engine* e, Handle h, int32_t* words);                     // flushes if needed; applies the negation

// Math.h routine scope: flushes at the outermost exit in eager mode (unless a batch is open)
struct RoutineScope {
    Engine* e;
    explicit RoutineScope(const TFheGateBootstrappingCloudKeySet* bk);
    ~RoutineScope();
};

// sample <-> handle
Handle sample_handle(Engine* e, const LweSample* s);  // borrowed (the sample keeps its reference); uploads host values
void sample_assign(LweSample* s, Handle h);           // the sample takes over one reference the caller owns

int recorder_mode();  // peba1_set_mode
void note_linear();  // statistics: a NOT/COPY/CONSTANT that cost nothing

}  // namespace peba1
