// tfhe_front.cpp — libpeba1_tfhe.so: the TFHE gate-bootstrapping API of include/tfhe/*.h as a recording front end
// of libpeba1_cuda (see recorder.hpp), plus host-side key generation and encryption/decryption.
//
// What stays on the host (north_star: "keygen still runs on the host"): parameter sets, LWE/TLWE/bootstrapping/
// keyswitch key generation, bootsSymEncrypt / bootsSymDecrypt (§3.4).  What goes to the GPU: every
// bootstrapped gate, batched per circuit level.  The integer bootstrapping key is uploaded once and transformed on
// the device.  No CPU implementation of the gates exists in this library: without a device the first gate aborts.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "../../include/tfhe/tfhe.h"
#include "chacha_rng.hpp"
#include "recorder.hpp"

namespace peba1 {

[[noreturn]] static void die(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "peba1: fatal: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    abort();  // upstream TFHE aborts on fatal errors too
}

// ------------------------------------------------------------------------------------------------ key material
}  // namespace peba1

struct peba1_key_material {
    TFheGateBootstrappingParameterSet params;
    std::vector<int32_t> lwe_key, tlwe_key, bk, ksk;
    peba1::Engine* engine = nullptr;
    TFheGateBootstrappingSecretKeySet* cloud_only_owner = nullptr;  // set for a cloud keyset loaded from a file
};

namespace peba1 {

struct Wire {
    int64_t slot = -1;   // device pool slot while the value is materialised
    int32_t refs = 0;
    int32_t level = 0;   // 0: available (materialised or queued for upload); >0: produced by a pending node
    int32_t next_free = -1;
};

struct Engine {
    // Everything the recorder knows lives in the engine of ONE key set (one GPU context): its wire table, its pending
    // levels, its statistics, behind its own lock.  Threads that drive different engines (one GPU each) never meet.
    std::recursive_mutex mu;
    std::vector<Wire> wires = std::vector<Wire>(2);  // wire 0 unused, wire 1 = constant false
    int32_t free_wire = -1;
    peba1_host_stats st{};
    int index = 0;
    peba1_key_material* km = nullptr;
    peba1_cuda_ctx* ctx = nullptr;          // null when a test backend is installed
    const peba1_test_backend* tb = nullptr;  // see peba1_host_set_test_backend (never set by the product)
    int words = 0;
    std::vector<std::vector<peba1_gate>> levels;  // pending nodes; in[]/out hold WIRE ids (or constant bits)
    std::vector<int64_t> up_wires;
    std::vector<int32_t> up_words;
    std::vector<int64_t> free_slots;
    int64_t next_slot = 0, capacity = 0;
    std::vector<peba1_gate> scratch;
    int routine_depth = 0;
};

// Process-wide: only the registry of engines (append-only) and the settings.  A handle names its engine in its upper bits
// ((engine index + 1) << 40 | wire << 1 | negated; the constants 2 and 3 belong to every engine), so samples can be
// released or fetched without a key argument and without a "current engine".
static std::recursive_mutex g_mu;        // guards g_engines, the settings below and the process-wide random generator
static std::vector<Engine*> g_engines;
static int g_device = -1;                // default device of engines created by threads that did not choose one
static thread_local int t_device = -1;   // peba1_host_set_device is per thread: one host thread per GPU in one process
static thread_local int t_batch_depth = 0;
static thread_local std::vector<Engine*> t_engines;  // engines this thread has used (what its flush / batch_end act on)
static std::atomic<bool> g_eager{true};
static std::atomic<uint64_t> g_linear{0};  // NOT / COPY / CONSTANT calls (no engine involved)
constexpr int kEngineShift = 40;
constexpr Handle kWireMask = (Handle(1) << kEngineShift) - 1;
static inline int64_t wire_of(Handle h) { return (int64_t)((h & kWireMask) >> 1); }
static inline Handle mk(const Engine* e, int64_t wire) { return (Handle(e->index + 1) << kEngineShift) | ((Handle)wire << 1); }
// lock-free lookup (called with an engine lock held): the table is append-only and published by the counter
constexpr int kMaxEngines = 4096;
static Engine* g_engine_tab[kMaxEngines];
static std::atomic<int> g_engine_count{0};
static Engine* engine_from(Handle h) {
    const int idx = (int)(h >> kEngineShift);
    if (idx <= 0 || idx > g_engine_count.load(std::memory_order_acquire)) die("handle %llx names no engine", (unsigned long long)h);
    return g_engine_tab[idx - 1];
}
static const peba1_test_backend* g_test_backend = nullptr;
static std::atomic<int> g_mode{0};  // peba1_set_mode: 0 faithful, 1 fold + dead-gate elimination, 2 = 1 + MUX-carry adders,
                        // 3 = 1 + 3-input full adders (XOR3 / MAJ3) and carry-save column compression (math_circuits.cpp)

// The four engine operations the recorder needs.  The product path is libpeba1_cuda; the table exists only so the
// CPU test-suite can observe the schedule (tests/ installs a bit-level simulator).
static int be_reserve(Engine* e, int64_t slots);
static int be_upload_slots(Engine* e, const int64_t* slots, int64_t count, const int32_t* words);
static int be_download(Engine* e, int64_t slot, int32_t* words);
static int be_gates(Engine* e, const peba1_gate* gates, int64_t count);

static int pick_device() {
    if (t_device >= 0) return t_device;
    if (g_device >= 0) return g_device;
    if (const char* s = getenv("PEBA1_DEVICE")) return atoi(s);
    if (const char* s = getenv("LOCAL_RANK")) return atoi(s);
    return 0;
}

#define CUCHK(e, call)                                                              \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if (rc_ != 0) die("%s failed (%d): %s", #call, rc_, peba1_cuda_last_error()); \
    } while (0)

Engine* engine_of(const TFheGateBootstrappingCloudKeySet* bk) {
    if (!bk || !bk->material) die("gate called with a null cloud key");
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    peba1_key_material* km = bk->material;
    if (!km->engine) {
        Engine* e = new Engine();
        e->km = km;
        e->index = (int)g_engines.size();
        e->words = km->params.n + 1;
        e->tb = g_test_backend;
        if (!e->tb) {
            peba1_cuda_params p{km->params.n, km->params.N, km->params.k, km->params.l, km->params.Bgbit,
                                km->params.ks_t, km->params.ks_basebit};
            CUCHK(e, peba1_cuda_create(pick_device(), &p, &e->ctx));
            if (!km->bk.empty()) {  // a placeholder keyset receives its device-format keys by NCCL broadcast instead
                CUCHK(e, peba1_cuda_load_bootstrap_key(e->ctx, km->bk.data()));
                CUCHK(e, peba1_cuda_load_keyswitch_key(e->ctx, km->ksk.data()));
            }
        }
        if (e->index >= kMaxEngines) die("more than %d key sets in one process", kMaxEngines);
        g_engine_tab[e->index] = e;
        g_engine_count.store(e->index + 1, std::memory_order_release);
        g_engines.push_back(e);
        km->engine = e;
    }
    Engine* e = km->engine;
    if (std::find(t_engines.begin(), t_engines.end(), e) == t_engines.end()) t_engines.push_back(e);
    return e;
}

static int be_reserve(Engine* e, int64_t slots) {
    return e->tb ? e->tb->reserve(e->tb->user, slots) : peba1_cuda_pool_reserve(e->ctx, slots);
}
static int be_upload_slots(Engine* e, const int64_t* slots, int64_t count, const int32_t* words) {
    return e->tb ? e->tb->upload_slots(e->tb->user, slots, count, words)
                 : peba1_cuda_upload_slots(e->ctx, slots, count, words);
}
static int be_download(Engine* e, int64_t slot, int32_t* words) {
    return e->tb ? e->tb->download(e->tb->user, slot, words) : peba1_cuda_download(e->ctx, slot, 1, words);
}
static int be_gates(Engine* e, const peba1_gate* gates, int64_t count) {
    return e->tb ? e->tb->gates(e->tb->user, gates, count) : peba1_cuda_gates(e->ctx, gates, count);
}

static int64_t new_wire(Engine* e, int level) {
    int64_t id;
    if (e->free_wire >= 0) {
        id = e->free_wire;
        e->free_wire = e->wires[id].next_free;
    } else {
        id = (int64_t)e->wires.size();
        e->wires.emplace_back();
    }
    Wire& w = e->wires[id];
    w.slot = -1;
    w.refs = 1;
    w.level = level;
    w.next_free = -1;
    return id;
}

static void unref_wire(Engine* e, int64_t id) {
    Wire& w = e->wires[id];
    if (--w.refs > 0) return;
    if (w.refs < 0) die("wire %lld released twice", (long long)id);
    if (w.slot >= 0) {
        e->free_slots.push_back(w.slot);
        e->st.live_slots--;
    }
    w.slot = -1;
    w.next_free = e->free_wire;
    e->free_wire = (int32_t)id;
}

void retain(Handle h) {
    if (h == 0 || is_const(h)) return;
    Engine* e = engine_from(h);
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    e->wires[wire_of(h)].refs++;
}
void release(Handle h) {
    if (h == 0 || is_const(h)) return;
    Engine* e = engine_from(h);
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    unref_wire(e, wire_of(h));
}

static int64_t alloc_slot(Engine* e) {
    e->st.live_slots++;
    if (e->st.live_slots > e->st.peak_slots) e->st.peak_slots = e->st.live_slots;
    if (!e->free_slots.empty()) {
        int64_t s = e->free_slots.back();
        e->free_slots.pop_back();
        return s;
    }
    return e->next_slot++;
}
static void ensure_capacity(Engine* e) {
    if (e->next_slot <= e->capacity) return;
    int64_t want = std::max<int64_t>(e->next_slot, std::max<int64_t>(4096, e->capacity * 2));
    CUCHK(e, be_reserve(e, want));
    e->capacity = want;
}

// fills operand q of a node from a handle; returns the operand's level
static int set_operand(Engine* e, peba1_gate& g, int q, Handle h) {
    if (h == 0) die("internal: host handle used as a gate operand");
    if (is_const(h)) {
        g.flags |= PEBA1_IN_CONST(q);
        g.in[q] = (int64_t)(h & 1);
        return 0;
    }
    if ((h >> kEngineShift) != (Handle)(e->index + 1))
        die("a gate mixes ciphertexts of two key sets (operand %d belongs to another engine)", q);
    g.in[q] = wire_of(h);
    if (h & 1) g.flags |= PEBA1_IN_NEG(q);
    Wire& w = e->wires[wire_of(h)];
    w.refs++;  // the pending node holds its operands
    return w.level;
}

static Handle push_node(Engine* e, peba1_gate& g, int level) {
    const int64_t out = new_wire(e, level);  // the node's own reference, dropped after execution
    e->wires[out].refs++;                    // the caller's reference
    g.out = out;
    if ((int)e->levels.size() < level) e->levels.resize(level);
    e->levels[level - 1].push_back(g);
    return mk(e, out);
}

// Boolean function of a 2-input gate (SURVEY.md §8a constants)
static inline int eval2(int op, int x, int y) {
    switch (op) {
        case PEBA1_OP_NAND: return 1 - (x & y);
        case PEBA1_OP_AND: return x & y;
        case PEBA1_OP_OR: return x | y;
        case PEBA1_OP_XOR: return x ^ y;
        case PEBA1_OP_XNOR: return 1 - (x ^ y);
        case PEBA1_OP_NOR: return 1 - (x | y);
        case PEBA1_OP_ANDNY: return (1 - x) & y;
        case PEBA1_OP_ANDYN: return x & (1 - y);
        case PEBA1_OP_ORNY: return (1 - x) | y;
        default: return x | (1 - y);  // ORYN
    }
}

static Handle record_gate_raw(Engine* e, int op, Handle a, Handle b) {
    peba1_gate g{};
    g.op = op;
    const int la = set_operand(e, g, 0, a), lb = set_operand(e, g, 1, b);
    e->st.gates++;
    e->st.pbs++;
    return push_node(e, g, 1 + std::max(la, lb));
}

// Fold mode (changes raw ciphertexts, never decrypted results): a gate or MUX whose operands
// involve fewer distinct variables than its arity is replaced by a constant, an alias (possibly negated: free), or a
// cheaper 2-input gate.  Operands are (variable wire, polarity) or constants.
struct FoldOperand { int64_t var; int neg; int cst; };  // var < 0: constant cst
static inline FoldOperand fold_operand(Handle h) {
    if (is_const(h)) return FoldOperand{-1, 0, (int)(h & 1)};
    return FoldOperand{wire_of(h), (int)(h & 1), 0};
}
static inline int fold_value(const FoldOperand& o, const int64_t* vars, const int* vals, int nvars) {
    if (o.var < 0) return o.cst;
    for (int q = 0; q < nvars; ++q)
        if (vars[q] == o.var) return vals[q] ^ o.neg;
    return 0;
}
// returns true and sets *out (one reference owned by the caller) when `nin` operands under `fn` reduce
template <class Fn>
static bool try_fold(Engine* e, const Handle* in, int nin, Fn fn, Handle* out) {
    FoldOperand o[3];
    int64_t vars[3];
    int nvars = 0;
    for (int q = 0; q < nin; ++q) {
        o[q] = fold_operand(in[q]);
        if (o[q].var < 0) continue;
        bool seen = false;
        for (int k = 0; k < nvars; ++k) seen |= vars[k] == o[q].var;
        if (!seen) vars[nvars++] = o[q].var;
    }
    if (nvars >= nin) return false;  // nothing to fold
    int table[4] = {0, 0, 0, 0};
    const int rows = 1 << nvars;
    for (int m = 0; m < rows; ++m) {
        const int vals[3] = {m & 1, (m >> 1) & 1, 0};
        int v[3];
        for (int q = 0; q < nin; ++q) v[q] = fold_value(o[q], vars, vals, nvars);
        table[m] = fn(v);
    }
    auto alias = [&](int64_t var, int neg) {
        const Handle h = mk(e, var) | (Handle)neg;
        retain(h);
        *out = h;
    };
    if (nvars == 0) { *out = table[0] ? kTrue : kFalse; e->st.folded++; return true; }
    if (nvars == 1) {
        if (table[0] == table[1]) *out = table[0] ? kTrue : kFalse;
        else alias(vars[0], table[0]);  // table = {0,1}: x ; {1,0}: NOT x
        e->st.folded++;
        return true;
    }
    // two variables x = vars[0] (bit 0 of m), y = vars[1] (bit 1 of m)
    const bool dep_x = table[0] != table[1] || table[2] != table[3];
    const bool dep_y = table[0] != table[2] || table[1] != table[3];
    if (!dep_x && !dep_y) { *out = table[0] ? kTrue : kFalse; e->st.folded++; return true; }
    if (!dep_y) { alias(vars[0], table[0]); e->st.folded++; return true; }
    if (!dep_x) { alias(vars[1], table[0]); e->st.folded++; return true; }
    for (int op = PEBA1_OP_NAND; op <= PEBA1_OP_ORYN; ++op) {
        bool ok = true;
        for (int m = 0; m < 4 && ok; ++m) ok = eval2(op, m & 1, (m >> 1) & 1) == table[m];
        if (ok) {
            *out = record_gate_raw(e, op, mk(e, vars[0]), mk(e, vars[1]));
            e->st.folded++;
            return true;
        }
    }
    return false;  // unreachable: the ten gate types cover every 2-variable function that depends on both
}

Handle record_gate(Engine* e, int op, Handle a, Handle b) {
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (g_mode >= 1) {
        const Handle in[2] = {a, b};
        Handle out;
        if (try_fold(e, in, 2, [op](const int* v) { return eval2(op, v[0], v[1]); }, &out)) return out;
    }
    return record_gate_raw(e, op, a, b);
}

Handle record_mux(Engine* e, Handle sel, Handle if1, Handle if0) {
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (g_mode >= 1) {
        const Handle in[3] = {sel, if1, if0};
        Handle out;
        if (try_fold(e, in, 3, [](const int* v) { return v[0] ? v[1] : v[2]; }, &out)) return out;
    }
    peba1_gate g{};
    g.op = PEBA1_OP_MUX;
    const int l0 = set_operand(e, g, 0, sel), l1 = set_operand(e, g, 1, if1), l2 = set_operand(e, g, 2, if0);
    e->st.gates++;
    e->st.pbs += 2;
    return push_node(e, g, 1 + std::max(l0, std::max(l1, l2)));
}

// 3-input gates in ONE bootstrap (PEBA1_OP_XOR3 / PEBA1_OP_MAJ3; not part of upstream's API, used by circuit mode 3):
// sum and carry of a full adder.  Constant, repeated or complementary operands fold like any other gate.
Handle record_gate3(Engine* e, int op, Handle a, Handle b, Handle c) {
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    const Handle in[3] = {a, b, c};
    Handle out;
    const bool maj = op == PEBA1_OP_MAJ3;
    if (try_fold(e, in, 3, [maj](const int* v) { return maj ? (v[0] + v[1] + v[2] >= 2) : (v[0] ^ v[1] ^ v[2]); }, &out))
        return out;
    peba1_gate g{};
    g.op = op;
    const int l0 = set_operand(e, g, 0, a), l1 = set_operand(e, g, 1, b), l2 = set_operand(e, g, 2, c);
    e->st.gates++;
    e->st.pbs += 1;
    return push_node(e, g, 1 + std::max(l0, std::max(l1, l2)));
}

int recorder_mode() { return g_mode; }

Handle wire_from_host(Engine* e, const int32_t* words) {
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    const int64_t id = new_wire(e, 0);  // reference owned by the upload queue
    e->wires[id].refs++;                // the caller's reference
    e->up_wires.push_back(id);
    e->up_words.insert(e->up_words.end(), words, words + e->words);
    return mk(e, id);
}

void flush(Engine* e) {
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (e->up_wires.empty() && e->levels.empty()) return;
    const auto t0 = std::chrono::steady_clock::now();
    // 1. pending uploads
    if (!e->up_wires.empty()) {
        std::vector<int64_t> slots(e->up_wires.size());
        for (size_t i = 0; i < slots.size(); ++i) slots[i] = e->wires[e->up_wires[i]].slot = alloc_slot(e);
        ensure_capacity(e);
        CUCHK(e, be_upload_slots(e, slots.data(), (int64_t)slots.size(), e->up_words.data()));
        e->st.uploads += slots.size();
        for (int64_t id : e->up_wires) unref_wire(e, id);
        e->up_wires.clear();
        e->up_words.clear();
    }
    // 1b. fold mode: dead-gate elimination.  A pending node whose output is referenced by nobody but itself (no
    // sample, no builder temporary, no live consumer) is dropped; walking the levels from the last to the first lets
    // the release of its operands cascade.
    if (g_mode >= 1) {
        for (size_t lv = e->levels.size(); lv-- > 0;) {
            auto& nodes = e->levels[lv];
            size_t keep = 0;
            for (size_t i = 0; i < nodes.size(); ++i) {
                const peba1_gate nd = nodes[i];
                if (e->wires[nd.out].refs == 1) {
                    const int nin = nd.op >= PEBA1_OP_MUX ? 3 : 2;
                    for (int q = 0; q < nin; ++q)
                        if (!(nd.flags & PEBA1_IN_CONST(q))) unref_wire(e, nd.in[q]);
                    unref_wire(e, nd.out);
                    e->st.dead++;
                } else {
                    nodes[keep++] = nd;
                }
            }
            nodes.resize(keep);
        }
    }
    // 2. one batched launch pair per level
    for (auto& nodes : e->levels) {
        if (nodes.empty()) continue;
        e->scratch.resize(nodes.size());
        for (size_t i = 0; i < nodes.size(); ++i) {
            const peba1_gate& nd = nodes[i];
            e->st.executed_gates++;
            e->st.executed_pbs += nd.op == PEBA1_OP_MUX ? 2 : 1;
            peba1_gate& g = e->scratch[i];
            g = nd;
            const int nin = nd.op >= PEBA1_OP_MUX ? 3 : 2;
            for (int q = 0; q < nin; ++q)
                if (!(nd.flags & PEBA1_IN_CONST(q))) g.in[q] = e->wires[nd.in[q]].slot;
            g.out = e->wires[nd.out].slot = alloc_slot(e);
        }
        ensure_capacity(e);
        CUCHK(e, be_gates(e, e->scratch.data(), (int64_t)nodes.size()));
        e->st.levels++;
        if (nodes.size() > e->st.max_level_width) e->st.max_level_width = nodes.size();
        for (const peba1_gate& nd : nodes) {
            const int nin = nd.op >= PEBA1_OP_MUX ? 3 : 2;
            e->wires[nd.out].level = 0;
            for (int q = 0; q < nin; ++q)
                if (!(nd.flags & PEBA1_IN_CONST(q))) unref_wire(e, nd.in[q]);
            unref_wire(e, nd.out);
        }
    }
    e->levels.clear();
    e->st.flushes++;
    e->st.flush_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void fetch(Engine* e, Handle h, int32_t* words) {
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (is_const(h)) {
        memset(words, 0, sizeof(int32_t) * (size_t)(e->words - 1));
        words[e->words - 1] = (h & 1) ? (1 << 29) : -(1 << 29);
        return;
    }
    Wire& w = e->wires[wire_of(h)];
    if (w.level > 0 || w.slot < 0) flush(e);
    if (e->wires[wire_of(h)].slot < 0) die("internal: wire %llu has no value after a flush", (unsigned long long)wire_of(h));
    CUCHK(e, be_download(e, e->wires[wire_of(h)].slot, words));
    e->st.downloads++;
    if (h & 1)
        for (int i = 0; i < e->words; ++i) words[i] = (int32_t)(0u - (uint32_t)words[i]);
}

RoutineScope::RoutineScope(const TFheGateBootstrappingCloudKeySet* bk) : e(engine_of(bk)) {
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    e->routine_depth++;
}
RoutineScope::~RoutineScope() {
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (--e->routine_depth == 0 && g_eager && t_batch_depth == 0) flush(e);
}

Handle sample_handle(Engine* e, const LweSample* s) {
    if (s->peba1_handle != 0) return s->peba1_handle;
    // host value: becomes a wire (queued upload); the sample keeps the reference so repeated uses upload once
    std::vector<int32_t> tmp((size_t)e->words);
    memcpy(tmp.data(), s->a, sizeof(int32_t) * (size_t)(e->words - 1));
    tmp[e->words - 1] = s->b;
    const Handle h = wire_from_host(e, tmp.data());
    const_cast<LweSample*>(s)->peba1_handle = h;
    return h;
}

void sample_assign(LweSample* s, Handle h) {
    const Handle old = s->peba1_handle;
    s->peba1_handle = h;
    release(old);
}

void note_linear() { g_linear.fetch_add(1, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------ randomness
// One process-wide generator like upstream — but a CSPRNG (ChaCha20, chacha_rng.hpp) keyed from the OS entropy pool at
// first use: upstream's default engine is unseeded and the reference never calls tfhe_random_generator_setSeed
// (SURVEY.md §8b), which would give every process the same, publicly derivable keys.  setSeed makes it deterministic.
static ChaChaRng g_rng_state;
static bool g_rng_keyed = false;
static ChaChaRng& global_rng() {  // callers hold g_mu
    if (!g_rng_keyed) {
        uint32_t key[8];
        if (!ChaChaRng::os_key(key)) die("cannot read the OS entropy pool (getrandom): refusing to generate predictable keys");
        g_rng_state = ChaChaRng(key, 0);
        g_rng_keyed = true;
    }
    return g_rng_state;
}

static inline int32_t double_to_torus32(double d) {  // fractional part of d, scaled to 2^32
    return (int32_t)(int64_t)((d - (double)(int64_t)d) * 4294967296.0);
}
template <class Rng>
static inline uint32_t gaussian32(Rng& rng, uint32_t message, double sigma) {
    std::normal_distribution<double> dist(0.0, sigma);
    return message + (uint32_t)double_to_torus32(dist(rng));
}
static inline uint32_t mod_switch_to_torus32(int32_t mu, int32_t msize) {
    const uint64_t interval = ((UINT64_C(1) << 63) / (uint64_t)msize) * 2;
    return (uint32_t)(((uint64_t)(int64_t)mu * interval) >> 32);
}

template <class Rng>
static void lwe_encrypt(Rng& rng, const int32_t* key, int n, uint32_t message, double sigma, int32_t* out) {
    uint32_t b = gaussian32(rng, message, sigma);
    for (int i = 0; i < n; ++i) {
        const uint32_t a = (uint32_t)rng();
        out[i] = (int32_t)a;
        if (key[i]) b += a;
    }
    out[n] = (int32_t)b;
}

// LWE key, ring key, keyswitch key (noise recentred, digit-0 rows noiseless zeros), n TGSW
// encryptions of the LWE key bits.  Rows are independent, so they are generated on all host cores with per-row
// generators derived from the master seed (deterministic for a given seed, independent of the thread count).
// All randomness of one key set comes from ONE 256-bit ChaCha20 key; the streams below (master, keyswitch noise, one
// per key row) are nonces under it.
static void generate_keys(peba1_key_material* km, const uint32_t key256[8]) {
    const TFheGateBootstrappingParameterSet& p = km->params;
    const int n = p.n, N = p.N, l = p.l, t = p.ks_t, base = 1 << p.ks_basebit, rows = (p.k + 1) * l;
    ChaChaRng master(key256, 0);
    km->lwe_key.resize(n);
    km->tlwe_key.resize(N);
    for (int i = 0; i < n; ++i) km->lwe_key[i] = (int32_t)(master() & 1);
    for (int i = 0; i < N; ++i) km->tlwe_key[i] = (int32_t)(master() & 1);
    km->ksk.assign((size_t)N * t * base * (n + 1), 0);
    km->bk.assign((size_t)n * rows * 2 * N, 0);

    // keyswitch noises first (their mean is removed, as upstream does), then the samples
    const size_t nks = (size_t)N * t * (base - 1);
    std::vector<double> noise(nks);
    {
        ChaChaRng r(key256, 1);
        std::normal_distribution<double> dist(0.0, p.ks_stdev);
        double mean = 0;
        for (double& x : noise) { x = dist(r); mean += x; }
        mean /= (double)nks;
        for (double& x : noise) x -= mean;
    }
    unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    auto parallel_for = [&](int count, auto&& body) {
        std::vector<std::thread> th;
        for (unsigned w = 0; w < nthreads; ++w)
            th.emplace_back([&, w] { for (int i = (int)w; i < count; i += (int)nthreads) body(i); });
        for (auto& x : th) x.join();
    };
    parallel_for(N, [&](int i) {
        ChaChaRng r(key256, 2 + (uint64_t)i);
        for (int j = 0; j < t; ++j)
            for (int h = 1; h < base; ++h) {
                int32_t* ct = km->ksk.data() + (((size_t)i * t + j) * base + h) * (size_t)(n + 1);
                const uint32_t mess = (uint32_t)(km->tlwe_key[i] * h) << (32 - (j + 1) * p.ks_basebit);
                uint32_t b = mess + (uint32_t)double_to_torus32(noise[((size_t)i * t + j) * (base - 1) + (h - 1)]);
                for (int q = 0; q < n; ++q) {
                    const uint32_t a = (uint32_t)r();
                    ct[q] = (int32_t)a;
                    if (km->lwe_key[q]) b += a;
                }
                ct[n] = (int32_t)b;
            }
    });
    // bootstrapping key: row r of BK_i is a TLWE encryption of 0 (b = a*s + e in the negacyclic ring) plus
    // s_i * 2^(32-(p+1)Bgbit) on coefficient 0 of the mask (rows 0..l-1) or of the body (rows l..2l-1)
    std::vector<int> ones;
    for (int m = 0; m < N; ++m) if (km->tlwe_key[m]) ones.push_back(m);
    parallel_for(n * rows, [&](int ir) {
        ChaChaRng r(key256, 2 + (uint64_t)N + (uint64_t)ir);
        uint32_t* a = (uint32_t*)km->bk.data() + ((size_t)ir * 2 + 0) * (size_t)N;
        uint32_t* b = (uint32_t*)km->bk.data() + ((size_t)ir * 2 + 1) * (size_t)N;
        for (int j = 0; j < N; ++j) b[j] = gaussian32(r, 0, p.bk_stdev);
        for (int j = 0; j < N; ++j) a[j] = (uint32_t)r();
        for (int m : ones) {  // b += X^m * a  (negacyclic)
            for (int j = 0; j < m; ++j) b[j] -= a[j - m + N];
            for (int j = m; j < N; ++j) b[j] += a[j - m];
        }
        const int i = ir / rows, row = ir % rows, bloc = row / l, pp = row % l;
        const uint32_t h = UINT32_C(1) << (32 - (pp + 1) * p.Bgbit);
        (bloc == 0 ? a : b)[0] += (uint32_t)km->lwe_key[i] * h;
    });
}

static TFheGateBootstrappingSecretKeySet* wrap_keyset(peba1_key_material* km) {
    auto* ks = new TFheGateBootstrappingSecretKeySet();
    ks->params = &km->params;  // a private copy: main.cpp deletes the parameter set before the ciphertexts
    ks->lwe_key = km->lwe_key.data();
    ks->tlwe_key = km->tlwe_key.data();
    ks->cloud.params = &km->params;
    ks->cloud.material = km;
    return ks;
}

static uint32_t phase_of(const TFheGateBootstrappingSecretKeySet* key, const int32_t* a, int32_t b) {
    const int n = key->params->n;
    uint32_t axs = 0;
    for (int i = 0; i < n; ++i)
        if (key->lwe_key[i]) axs += (uint32_t)a[i];
    return (uint32_t)b - axs;
}

}  // namespace peba1

using namespace peba1;

// ================================================================================================ C API
extern "C" {

TFheGateBootstrappingParameterSet* new_default_gate_bootstrapping_parameters(int32_t minimum_lambda) {
    if (minimum_lambda > 128) die("sorry, for now the parameters are only implemented for 80 and 128 bits of security");
    auto* p = new TFheGateBootstrappingParameterSet();
    p->ks_t = 8;
    p->ks_basebit = 2;
    p->N = 1024;
    p->k = 1;
    p->minimum_lambda = minimum_lambda;
    if (minimum_lambda > 80) {
        p->n = 630; p->l = 3; p->Bgbit = 7;
        p->ks_stdev = ldexp(1.0, -15); p->bk_stdev = ldexp(1.0, -25); p->max_stdev = 0.012467;
    } else {
        p->n = 500; p->l = 2; p->Bgbit = 10;
        p->ks_stdev = 2.44e-5; p->bk_stdev = 7.18e-9; p->max_stdev = 0.012467;
    }
    return p;
}
void delete_gate_bootstrapping_parameters(TFheGateBootstrappingParameterSet* params) { delete params; }

TFheGateBootstrappingSecretKeySet* peba1_new_secret_keyset_seeded(const TFheGateBootstrappingParameterSet* params, uint64_t seed) {
    if (!params) die("null parameter set");
    auto* km = new peba1_key_material();
    km->params = *params;
    uint32_t key[8];
    ChaChaRng::expand_seed(seed, key);  // deterministic: tests and benchmarks only (include/peba1_host.h)
    generate_keys(km, key);
    return wrap_keyset(km);
}
TFheGateBootstrappingSecretKeySet* new_random_gate_bootstrapping_secret_keyset(const TFheGateBootstrappingParameterSet* params) {
    if (!params) die("null parameter set");
    uint32_t key[8];
    {
        std::lock_guard<std::recursive_mutex> lk(g_mu);
        ChaChaRng& g = global_rng();
        for (int i = 0; i < 4; ++i) { const uint64_t v = g(); key[2 * i] = (uint32_t)v; key[2 * i + 1] = (uint32_t)(v >> 32); }
    }
    auto* km = new peba1_key_material();
    km->params = *params;
    generate_keys(km, key);
    return wrap_keyset(km);
}
TFheGateBootstrappingSecretKeySet* peba1_new_keyset_from_arrays(const TFheGateBootstrappingParameterSet* params, const int32_t* lwe_key,
                                                                const int32_t* tlwe_key, const int32_t* bk, const int32_t* ksk) {
    if (!params || !bk || !ksk) die("null argument");
    auto* km = new peba1_key_material();
    km->params = *params;
    const TFheGateBootstrappingParameterSet& p = km->params;
    km->lwe_key.assign((size_t)p.n, 0);
    km->tlwe_key.assign((size_t)p.N, 0);
    if (lwe_key) memcpy(km->lwe_key.data(), lwe_key, sizeof(int32_t) * (size_t)p.n);
    if (tlwe_key) memcpy(km->tlwe_key.data(), tlwe_key, sizeof(int32_t) * (size_t)p.N);
    km->bk.assign(bk, bk + (size_t)p.n * (p.k + 1) * p.l * 2 * p.N);
    km->ksk.assign(ksk, ksk + (size_t)p.N * p.ks_t * (1 << p.ks_basebit) * (p.n + 1));
    return wrap_keyset(km);
}
TFheGateBootstrappingSecretKeySet* peba1_new_placeholder_keyset(const TFheGateBootstrappingParameterSet* params) {
    if (!params) die("null parameter set");
    auto* km = new peba1_key_material();
    km->params = *params;
    km->lwe_key.assign((size_t)params->n, 0);
    km->tlwe_key.assign((size_t)params->N, 0);
    return wrap_keyset(km);
}
void delete_gate_bootstrapping_secret_keyset(TFheGateBootstrappingSecretKeySet* keyset) {
    if (!keyset) return;
    peba1_key_material* km = keyset->cloud.material;
    // the engine (and the wires of still-allocated ciphertexts) outlives the keyset: main.cpp:600-607 deletes the
    // keyset before the ciphertext arrays.  Pending work is executed, host key copies are dropped.
    if (km->engine) flush(km->engine);
    km->bk.clear(); km->bk.shrink_to_fit();
    km->ksk.clear(); km->ksk.shrink_to_fit();
    delete keyset;
}
// A cloud keyset is normally embedded in its secret keyset (src/main.cpp:23 takes its address; :601 is commented out), so
// there is nothing to free — except for one loaded from a file (tfhe_io), which owns a key-less wrapper.
void delete_gate_bootstrapping_cloud_keyset(TFheGateBootstrappingCloudKeySet* cloud) {
    if (cloud && cloud->material && cloud->material->cloud_only_owner &&
        &cloud->material->cloud_only_owner->cloud == cloud)
        delete_gate_bootstrapping_secret_keyset(cloud->material->cloud_only_owner);
}
TFheGateBootstrappingCloudKeySet* peba1_new_cloud_keyset_from_arrays(const TFheGateBootstrappingParameterSet* params,
                                                                     const int32_t* bk, const int32_t* ksk) {
    TFheGateBootstrappingSecretKeySet* ks = peba1_new_keyset_from_arrays(params, nullptr, nullptr, bk, ksk);
    ks->cloud.material->cloud_only_owner = ks;
    ks->lwe_key = nullptr;   // a cloud keyset cannot encrypt or decrypt
    ks->tlwe_key = nullptr;
    return &ks->cloud;
}

LweSample* new_gate_bootstrapping_ciphertext_array(int32_t nbelems, const TFheGateBootstrappingParameterSet* params) {
    if (nbelems < 0 || !params) die("bad ciphertext array request");
    const size_t n = (size_t)params->n;
    const size_t head = ((sizeof(LweSample) * (size_t)nbelems + 63) / 64) * 64;
    char* block = (char*)malloc(head + sizeof(int32_t) * n * (size_t)nbelems + 64);
    if (!block) die("out of memory");
    LweSample* arr = (LweSample*)block;
    int32_t* words = (int32_t*)(block + head);
    for (int32_t i = 0; i < nbelems; ++i) {
        arr[i].a = words + n * (size_t)i;
        arr[i].b = 0;
        arr[i].current_variance = 0.0;
        arr[i].peba1_handle = kFalse;  // never-written ciphertexts read as trivial false
        arr[i].peba1_n = params->n;
        arr[i].peba1_pad = 0;
    }
    return arr;
}
LweSample* new_gate_bootstrapping_ciphertext(const TFheGateBootstrappingParameterSet* params) {
    return new_gate_bootstrapping_ciphertext_array(1, params);
}
void delete_gate_bootstrapping_ciphertext_array(int32_t nbelems, LweSample* samples) {
    if (!samples) return;
    for (int32_t i = 0; i < nbelems; ++i) release(samples[i].peba1_handle);
    free(samples);
}
void delete_gate_bootstrapping_ciphertext(LweSample* sample) { delete_gate_bootstrapping_ciphertext_array(1, sample); }

void bootsSymEncrypt(LweSample* result, int32_t message, const TFheGateBootstrappingSecretKeySet* key) {
    const int n = key->params->n;
    std::vector<int32_t> w((size_t)n + 1);
    const uint32_t mu = mod_switch_to_torus32(1, 8);
    {
        std::lock_guard<std::recursive_mutex> lk(g_mu);  // the process-wide generator only
        lwe_encrypt(global_rng(), key->lwe_key, n, message ? mu : 0u - mu, key->params->ks_stdev, w.data());
    }
    peba1_sample_set(result, w.data());
}

int32_t bootsSymDecrypt(const LweSample* sample, const TFheGateBootstrappingSecretKeySet* key) {
    const Handle h = sample->peba1_handle;
    if (h == 0) return (int32_t)phase_of(key, sample->a, sample->b) > 0 ? 1 : 0;
    if (is_const(h)) return (int32_t)(h & 1);
    Engine* e = engine_of(&key->cloud);
    std::vector<int32_t> w((size_t)e->words);
    fetch(e, h, w.data());
    // cache the words in the sample's host storage (the handle stays authoritative)
    LweSample* s = const_cast<LweSample*>(sample);
    memcpy(s->a, w.data(), sizeof(int32_t) * (size_t)(e->words - 1));
    s->b = w[e->words - 1];
    return (int32_t)phase_of(key, s->a, s->b) > 0 ? 1 : 0;
}

void bootsCONSTANT(LweSample* result, int32_t value, const TFheGateBootstrappingCloudKeySet*) {
    note_linear();
    sample_assign(result, value ? kTrue : kFalse);
}
void bootsNOT(LweSample* result, const LweSample* ca, const TFheGateBootstrappingCloudKeySet* bk) {
    note_linear();
    const Handle h = sample_handle(engine_of(bk), ca);
    retain(h);
    sample_assign(result, negate(h));
}
void bootsCOPY(LweSample* result, const LweSample* ca, const TFheGateBootstrappingCloudKeySet* bk) {
    note_linear();
    if (result == ca) return;
    const Handle h = sample_handle(engine_of(bk), ca);
    retain(h);
    sample_assign(result, h);
}

static void gate2(int op, LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk) {
    Engine* e = engine_of(bk);
    const Handle a = sample_handle(e, ca), b = sample_handle(e, cb);
    sample_assign(result, record_gate(e, op, a, b));  // operands are held by the node before the output is rebound
}
void bootsNAND(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_NAND, r, a, b, bk); }
void bootsOR(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_OR, r, a, b, bk); }
void bootsAND(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_AND, r, a, b, bk); }
void bootsXOR(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_XOR, r, a, b, bk); }
void bootsXNOR(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_XNOR, r, a, b, bk); }
void bootsNOR(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_NOR, r, a, b, bk); }
void bootsANDNY(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_ANDNY, r, a, b, bk); }
void bootsANDYN(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_ANDYN, r, a, b, bk); }
void bootsORNY(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_ORNY, r, a, b, bk); }
void bootsORYN(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { gate2(PEBA1_OP_ORYN, r, a, b, bk); }
void bootsMUX(LweSample* r, const LweSample* a, const LweSample* b, const LweSample* c, const TFheGateBootstrappingCloudKeySet* bk) {
    Engine* e = engine_of(bk);
    const Handle ha = sample_handle(e, a), hb = sample_handle(e, b), hc = sample_handle(e, c);
    sample_assign(r, record_mux(e, ha, hb, hc));
}

void tfhe_random_generator_setSeed(uint32_t* values, int32_t size) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    // upstream's hook: from here on the process-wide generator is deterministic (key = the values, absorbed 8 words at a
    // time).  The reference never calls it; tests do.
    uint32_t key[8] = {0x5eba1u, 0, 0, 0, 0, 0, 0, 0};
    for (int32_t i = 0; i < size; ++i) key[i & 7] = key[i & 7] * 0x9E3779B1u + values[i] + (uint32_t)i;
    g_rng_state = ChaChaRng(key, 0);
    g_rng_keyed = true;
}

// ------------------------------------------------------------------------------------------------ extensions
// Synthetic inputs generated on the device: arrays[k][i] <- fresh encryption of
// bits[k*nbelems + i], one kernel launch for the whole list, without the 2.5 KB per bit crossing PCIe.  With a test
// backend installed the words are made on the host and uploaded instead.
void peba1_encrypt_array_list_device(LweSample* const* arrays, int32_t narrays, int32_t nbelems, const int32_t* bits,
                                     const TFheGateBootstrappingSecretKeySet* key, uint64_t seed) {
    if (narrays <= 0 || nbelems <= 0) return;
    if (!arrays || !bits || !key || !key->lwe_key) die("peba1_encrypt_array_list_device: bad arguments");
    Engine* e = engine_of(&key->cloud);
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    const int64_t total = (int64_t)narrays * nbelems;
    std::vector<int64_t> slots((size_t)total);
    for (int32_t k = 0; k < narrays; ++k)
        for (int32_t i = 0; i < nbelems; ++i) {
            const int64_t id = new_wire(e, 0);  // the sample owns the only reference
            slots[(size_t)k * nbelems + i] = e->wires[id].slot = alloc_slot(e);
            sample_assign(&arrays[k][i], mk(e, id));
        }
    ensure_capacity(e);
    if (e->tb) {
        std::vector<int32_t> w((size_t)total * (size_t)e->words);
        peba1_encrypt_bits(key, bits, total, seed, w.data());
        CUCHK(e, be_upload_slots(e, slots.data(), total, w.data()));
    } else {
        CUCHK(e, peba1_cuda_encrypt_bits(e->ctx, key->lwe_key, bits, slots.data(), 0, total, 0, seed,
                                         key->params->ks_stdev));
    }
    e->st.device_encryptions += (uint64_t)total;
}
// arrays[k][i] <- sample (k*nbelems + i) of a dense DEVICE buffer ((n+1) words each), e.g. what an NCCL broadcast or
// scatter delivered: the samples become pool-resident wires by one device-side scatter, nothing crosses PCIe.
void peba1_sample_set_list_device(LweSample* const* arrays, int32_t narrays, int32_t nbelems, const void* device_words,
                                  const TFheGateBootstrappingCloudKeySet* bk) {
    if (narrays <= 0 || nbelems <= 0) return;
    if (!arrays || !device_words) die("peba1_sample_set_list_device: bad arguments");
    Engine* e = engine_of(bk);
    if (e->tb) die("peba1_sample_set_list_device needs the CUDA engine (a test backend holds no device memory)");
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    const int64_t total = (int64_t)narrays * nbelems;
    std::vector<int64_t> slots((size_t)total);
    for (int32_t k = 0; k < narrays; ++k)
        for (int32_t i = 0; i < nbelems; ++i) {
            const int64_t id = new_wire(e, 0);  // the sample owns the only reference
            slots[(size_t)k * nbelems + i] = e->wires[id].slot = alloc_slot(e);
            sample_assign(&arrays[k][i], mk(e, id));
        }
    ensure_capacity(e);
    CUCHK(e, peba1_cuda_upload_slots_device(e->ctx, slots.data(), 0, total, (const int32_t*)device_words));
    e->st.uploads += (uint64_t)total;
}
// the device pool row of a sample that already lives there (-1 otherwise; flushes pending work first): lets a caller
// gather result ciphertexts on the device, e.g. as the send buffer of an NCCL all-gather
int64_t peba1_sample_device_slot(const LweSample* sample, const TFheGateBootstrappingCloudKeySet* bk) {
    Engine* e = engine_of(bk);
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    const Handle h = sample->peba1_handle;
    if (h == 0 || is_const(h) || (h & 1)) return -1;  // host value, constant, or a negated alias: no row of its own
    if (e->wires[wire_of(h)].level > 0 || e->wires[wire_of(h)].slot < 0) flush(e);
    return e->wires[wire_of(h)].slot;
}
void peba1_encrypt_array_device(LweSample* array, int32_t nbelems, const int32_t* bits,
                                const TFheGateBootstrappingSecretKeySet* key, uint64_t seed) {
    peba1_encrypt_array_list_device(&array, 1, nbelems, bits, key, seed);
}

void peba1_host_set_device(int device) {  // this thread's engines; also the default of threads that never call it
    t_device = device;
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g_device = device;
}
void peba1_host_set_test_backend(const peba1_test_backend* backend) { g_test_backend = backend; }
static std::vector<Engine*> all_engines() {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    return g_engines;
}
// flush / batch scope act on the engines THIS THREAD has used (one host thread per GPU never touches another's)
void peba1_flush(void) {
    for (Engine* e : t_engines) flush(e);
}
void peba1_batch_begin(void) { t_batch_depth++; }
void peba1_batch_end(void) {
    if (t_batch_depth > 0 && --t_batch_depth == 0) peba1_flush();
}
void peba1_set_eager(int eager) { g_eager = eager != 0; }
void peba1_set_mode(int mode) {
    for (Engine* e : all_engines()) flush(e);  // pending work was recorded under the old mode
    g_mode = mode < 0 ? 0 : mode > 3 ? 3 : mode;
}
int peba1_get_mode(void) { return g_mode; }
void peba1_get_stats(peba1_host_stats* out) {  // summed over all engines of the process
    if (!out) return;
    peba1_host_stats t{};
    for (Engine* e : all_engines()) {
        std::lock_guard<std::recursive_mutex> lk(e->mu);
        const peba1_host_stats& s = e->st;
        t.gates += s.gates; t.pbs += s.pbs; t.levels += s.levels; t.flushes += s.flushes;
        t.max_level_width = std::max(t.max_level_width, s.max_level_width);
        t.uploads += s.uploads; t.downloads += s.downloads; t.live_slots += s.live_slots; t.peak_slots += s.peak_slots;
        t.flush_seconds += s.flush_seconds; t.executed_gates += s.executed_gates; t.executed_pbs += s.executed_pbs;
        t.folded += s.folded; t.dead += s.dead; t.device_encryptions += s.device_encryptions;
    }
    t.linear = g_linear.load(std::memory_order_relaxed);
    *out = t;
}
void peba1_reset_stats(void) {
    for (Engine* e : all_engines()) {
        std::lock_guard<std::recursive_mutex> lk(e->mu);
        const uint64_t live = e->st.live_slots;
        e->st = peba1_host_stats{};
        e->st.live_slots = e->st.peak_slots = live;
    }
    g_linear = 0;
}

int32_t peba1_sample_words(const TFheGateBootstrappingParameterSet* params) { return params->n + 1; }
void peba1_sample_get(const LweSample* sample, const TFheGateBootstrappingCloudKeySet* bk, int32_t* words) {
    const int n = bk->params->n;
    const Handle h = sample->peba1_handle;
    if (h == 0) {
        memcpy(words, sample->a, sizeof(int32_t) * (size_t)n);
        words[n] = sample->b;
        return;
    }
    if (is_const(h)) {
        memset(words, 0, sizeof(int32_t) * (size_t)n);
        words[n] = (h & 1) ? (1 << 29) : -(1 << 29);
        return;
    }
    fetch(engine_of(bk), h, words);
}
void peba1_sample_get_words(const LweSample* sample, int32_t* words) {
    const int n = sample->peba1_n;
    const Handle h = sample->peba1_handle;
    if (h == 0) {
        memcpy(words, sample->a, sizeof(int32_t) * (size_t)n);
        words[n] = sample->b;
    } else if (is_const(h)) {
        memset(words, 0, sizeof(int32_t) * (size_t)n);
        words[n] = (h & 1) ? (1 << 29) : -(1 << 29);
    } else {
        fetch(engine_from(h), h, words);
    }
}
void peba1_sample_set(LweSample* sample, const int32_t* words) {
    // words: a[0..n-1], b
    const Handle old = sample->peba1_handle;
    sample->peba1_handle = 0;
    release(old);
    const int n = sample->peba1_n;
    memcpy(sample->a, words, sizeof(int32_t) * (size_t)n);
    sample->b = words[n];
}
LweSample* peba1_sample_at(LweSample* array, int32_t index) { return &array[index]; }

const int32_t* peba1_keyset_bk(const TFheGateBootstrappingCloudKeySet* bk) { return bk->material->bk.empty() ? nullptr : bk->material->bk.data(); }
const int32_t* peba1_keyset_ksk(const TFheGateBootstrappingCloudKeySet* bk) { return bk->material->ksk.empty() ? nullptr : bk->material->ksk.data(); }

void peba1_encrypt_bits(const TFheGateBootstrappingSecretKeySet* key, const int32_t* bits, int64_t count, uint64_t seed, int32_t* words) {
    const int n = key->params->n;
    const uint32_t mu = mod_switch_to_torus32(1, 8);
    ChaChaRng r = ChaChaRng::from_seed(seed, 0x656e63);  // explicit seed: deterministic by request
    for (int64_t i = 0; i < count; ++i)
        lwe_encrypt(r, key->lwe_key, n, bits[i] ? mu : 0u - mu, key->params->ks_stdev, words + (size_t)i * (n + 1));
}
void peba1_decrypt_words(const TFheGateBootstrappingSecretKeySet* key, const int32_t* words, int64_t count, int32_t* bits) {
    const int n = key->params->n;
    for (int64_t i = 0; i < count; ++i) {
        const int32_t* w = words + (size_t)i * (n + 1);
        bits[i] = (int32_t)phase_of(key, w, w[n]) > 0 ? 1 : 0;
    }
}
void* peba1_cuda_context_of(const TFheGateBootstrappingCloudKeySet* bk) { return engine_of(bk)->ctx; }

void peba1_host_chacha_block(const uint32_t key[8], uint64_t block, uint64_t stream, uint64_t out[8]) {
    ChaChaRng r(key, stream);
    r.seek(block);
    for (int i = 0; i < 8; ++i) out[i] = r();
}

}  // extern "C"
