/*
 * oracle/tfhe_api_cpu.cpp — CPU ORACLE (test infrastructure, NOT the product).
 *
 * The TFHE gate API that the reference's circuits call (include/tfhe/*.h), evaluated on the CPU oracle
 * (oracle/tfhe_oracle.c).  It exists so that the reference's OWN src/Math.cpp — compiled unchanged from where it lies
 * under /root/reference by `make -C oracle ref` — can run here and serve as the circuit-level oracle:
 *
 *     reference Math.cpp (Math.cpp:27-417)  ->  boots* below  ->  oracle_gate_batch (FP64-FFT form == exact form,
 *                                                                  tests/test_oracle.py)
 *
 * The gate calls are the reference's, in the reference's order, on the reference's temporaries; this file only decides
 * WHEN a gate's arithmetic runs: a call appends a node to a list, and the list is evaluated (all host threads, gates
 * whose operands are ready first) when somebody needs a value (bootsSymDecrypt, refapi_sample_get, refapi_flush).
 * Gate outputs do not depend on evaluation order (the gate path uses no randomness), so the words
 * are those of a call-by-call evaluation; `refapi_set_eager(1)` evaluates call by call to check exactly that.
 *
 * Conventions shared with the product so that raw words can be compared:
 *   - a ciphertext that was never written reads as the trivial encryption of false (quirk Q2; upstream leaves it
 *     uninitialised, the reference reads such samples at Math.cpp:360 `tmp_sum[j] <- result[j]`);
 *   - arrays are over-allocated by REFAPI_PAD samples: the reference writes one sample past some of its arrays
 *     (Math.cpp:394-401: bootsSUBNbit writes result[bitsize] of a bitsize-long array).
 *
 * Nothing under peba1_b200/ or include/ links, loads or mentions this file.
 */
#include <tfhe/tfhe.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "tfhe_oracle.h"

#define REFAPI_PAD 8

struct peba1_key_material {
    oracle_keys* keys;
    bool owned;
};

namespace {

enum { OPX_NOT = 100, OPX_COPY = 101 };

struct Value {
    int32_t* words = nullptr; /* n+1 words once evaluated */
    int32_t level = 0;        /* circuit depth: 0 for inputs and constants */
    int32_t readers = 0;      /* pending nodes that still read it */
};

struct Node {
    int32_t op;
    int64_t in[3];
    int64_t out;
};

struct State {
    std::recursive_mutex mu;
    std::vector<Value> values;            /* index = handle - 1 */
    std::vector<Node> pending;
    std::unordered_map<LweSample*, int32_t> arrays; /* base pointer -> allocated length incl. padding */
    const oracle_keys* keys = nullptr;    /* the keyset the pending gates were recorded under */
    int words = 0;
    int threads = 0;
    int eager = 0;
    uint64_t rng = 0x5eba1c0ffeeULL;
    uint64_t seed_counter = 1;
    /* statistics */
    uint64_t gates = 0, pbs = 0, linear = 0, levels = 0, flushes = 0, max_width = 0;
    std::vector<uint64_t> level_widths;   /* of the last flush */
} S;

void die(const char* what) {
    fprintf(stderr, "tfhe_api_cpu (oracle): %s\n", what);
    abort();
}

int64_t new_value(int32_t level) {
    S.values.emplace_back();
    S.values.back().level = level;
    return (int64_t)S.values.size(); /* handle */
}

Value& val(int64_t handle) { return S.values[(size_t)handle - 1]; }

int32_t* alloc_words() { return (int32_t*)malloc(sizeof(int32_t) * (size_t)S.words); }

void bind_keys(const TFheGateBootstrappingCloudKeySet* bk) {
    if (!bk || !bk->material || !bk->material->keys) die("gate called without key material");
    if (S.keys != bk->material->keys) {
        if (!S.pending.empty()) die("two keysets interleaved without a flush");
        S.keys = bk->material->keys;
        S.words = oracle_keys_params(S.keys)->n + 1;
    }
}

/* handle of the value a sample currently names; a never-written sample becomes trivial false on first read (Q2) */
int64_t read_handle(const LweSample* s) {
    if (s->peba1_handle) return (int64_t)s->peba1_handle;
    if (!S.words) die("sample read before any keyset was seen");
    int64_t h = new_value(0);
    val(h).words = alloc_words();
    oracle_trivial(oracle_keys_params(S.keys), 0, val(h).words);
    const_cast<LweSample*>(s)->peba1_handle = (uint64_t)h;
    return h;
}

void flush_locked();

void record(int op, LweSample* out, const LweSample* a, const LweSample* b, const LweSample* c,
            const TFheGateBootstrappingCloudKeySet* bk) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    bind_keys(bk);
    Node n;
    n.op = op;
    n.in[0] = read_handle(a);
    n.in[1] = b ? read_handle(b) : 0;
    n.in[2] = c ? read_handle(c) : 0;
    int32_t lvl = 0;
    for (int q = 0; q < 3; ++q)
        if (n.in[q]) {
            lvl = std::max(lvl, val(n.in[q]).level);
            val(n.in[q]).readers++;
        }
    const bool lin = op >= OPX_NOT;
    n.out = new_value(lin ? lvl : lvl + 1);
    out->peba1_handle = (uint64_t)n.out; /* after the operands were read: out may alias an input (Math.cpp:272) */
    S.pending.push_back(n);
    if (lin) S.linear++;
    else { S.gates++; S.pbs += (op == ORACLE_OP_MUX) ? 2 : 1; }
    if (S.eager) flush_locked();
}

void flush_locked() {
    if (S.pending.empty()) return;
    const int threads = S.threads > 0 ? S.threads : (int)std::max(1u, std::thread::hardware_concurrency());
    const size_t W = (size_t)S.words;
    /* values some live sample still names must survive the flush */
    std::unordered_set<int64_t> held;
    for (auto& kv : S.arrays)
        for (int32_t i = 0; i < kv.second; ++i)
            if (kv.first[i].peba1_handle) held.insert((int64_t)kv.first[i].peba1_handle);
    /* nodes by level (a linear node sits on its operand's level and runs before that level's consumers) */
    int32_t maxlvl = 0;
    for (auto& n : S.pending) maxlvl = std::max(maxlvl, val(n.out).level);
    std::vector<std::vector<const Node*>> boot((size_t)maxlvl + 1), lin((size_t)maxlvl + 1);
    for (auto& n : S.pending) (n.op >= OPX_NOT ? lin : boot)[(size_t)val(n.out).level].push_back(&n);
    S.level_widths.assign((size_t)maxlvl + 1, 0);
    std::vector<int32_t> ca, cb, cc, out;
    auto release = [&](int64_t h) {
        Value& v = val(h);
        if (--v.readers == 0 && !held.count(h)) { free(v.words); v.words = nullptr; }
    };
    for (int32_t L = 0; L <= maxlvl; ++L) {
        /* bootstrapped gates of this level, one oracle batch per opcode */
        std::vector<const Node*>& B = boot[(size_t)L];
        S.level_widths[(size_t)L] = B.size();
        if (!B.empty()) { S.levels++; S.max_width = std::max<uint64_t>(S.max_width, B.size()); }
        std::stable_sort(B.begin(), B.end(), [](const Node* x, const Node* y) { return x->op < y->op; });
        for (size_t i = 0; i < B.size();) {
            size_t j = i;
            while (j < B.size() && B[j]->op == B[i]->op) ++j;
            const size_t cnt = j - i;
            const bool mux = B[i]->op == ORACLE_OP_MUX;
            ca.resize(cnt * W); cb.resize(cnt * W); out.resize(cnt * W);
            if (mux) cc.resize(cnt * W);
            for (size_t g = 0; g < cnt; ++g) {
                const Node* n = B[i + g];
                for (int q = 0; q < (mux ? 3 : 2); ++q) {
                    const int32_t* src = val(n->in[q]).words;
                    if (!src) die("operand not evaluated (scheduler bug)");
                    memcpy((q == 0 ? ca : q == 1 ? cb : cc).data() + g * W, src, W * sizeof(int32_t));
                }
            }
            if (oracle_gate_batch(S.keys, ORACLE_FFT, B[i]->op, ca.data(), cb.data(), mux ? cc.data() : nullptr,
                                  out.data(), (int64_t)cnt, threads) != 0)
                die("oracle_gate_batch failed");
            for (size_t g = 0; g < cnt; ++g) {
                Value& v = val(B[i + g]->out);
                v.words = alloc_words();
                memcpy(v.words, out.data() + g * W, W * sizeof(int32_t));
            }
            i = j;
        }
        /* NOT / COPY whose operand sits on this level (chains of them are in recording order) */
        for (const Node* n : lin[(size_t)L]) {
            const int32_t* src = val(n->in[0]).words;
            if (!src) die("linear operand not evaluated (scheduler bug)");
            Value& v = val(n->out);
            v.words = alloc_words();
            if (n->op == OPX_NOT) oracle_not(oracle_keys_params(S.keys), src, v.words);
            else memcpy(v.words, src, W * sizeof(int32_t));
        }
        for (const Node* n : B)
            for (int q = 0; q < 3; ++q) if (n->in[q]) release(n->in[q]);
        for (const Node* n : lin[(size_t)L]) release(n->in[0]);
    }
    /* results nobody names any more */
    for (auto& n : S.pending) {
        Value& v = val(n.out);
        if (v.readers == 0 && !held.count(n.out) && v.words) { free(v.words); v.words = nullptr; }
        v.level = 0; /* depth is counted per flush, like the product's launch pairs */
    }
    for (int64_t h : held) val(h).level = 0;
    S.pending.clear();
    S.flushes++;
}

const int32_t* words_of(const LweSample* s) {
    int64_t h = read_handle(s);
    if (!val(h).words) flush_locked();
    if (!val(h).words) die("value lost (it was released while a sample still named it)");
    return val(h).words;
}

} // namespace

extern "C" {

/* ---- parameters and keys (main.cpp:21-23, 600-602) ---- */
TFheGateBootstrappingParameterSet* new_default_gate_bootstrapping_parameters(int32_t minimum_lambda) {
    oracle_params op;
    if (oracle_default_params(minimum_lambda, &op) != 0) die("no parameter set for this security level");
    auto* p = (TFheGateBootstrappingParameterSet*)calloc(1, sizeof(TFheGateBootstrappingParameterSet));
    p->ks_t = op.ks_t; p->ks_basebit = op.ks_basebit; p->n = op.n; p->N = op.N; p->k = op.k; p->l = op.l;
    p->Bgbit = op.Bgbit; p->ks_stdev = op.ks_stdev; p->bk_stdev = op.bk_stdev; p->max_stdev = 0.012467;
    p->minimum_lambda = minimum_lambda;
    return p;
}
void delete_gate_bootstrapping_parameters(TFheGateBootstrappingParameterSet* p) { free(p); }

static oracle_params to_oracle_params(const TFheGateBootstrappingParameterSet* p) {
    oracle_params op;
    op.n = p->n; op.N = p->N; op.k = p->k; op.l = p->l; op.Bgbit = p->Bgbit; op.ks_t = p->ks_t;
    op.ks_basebit = p->ks_basebit; op.ks_stdev = p->ks_stdev; op.bk_stdev = p->bk_stdev;
    return op;
}

static TFheGateBootstrappingSecretKeySet* wrap_keys(const TFheGateBootstrappingParameterSet* params, oracle_keys* k,
                                                    bool owned) {
    auto* sk = (TFheGateBootstrappingSecretKeySet*)calloc(1, sizeof(TFheGateBootstrappingSecretKeySet));
    sk->params = params;
    sk->lwe_key = oracle_lwe_key(k);
    sk->tlwe_key = oracle_tlwe_key(k);
    sk->cloud.params = params;
    sk->cloud.material = new peba1_key_material{k, owned};
    return sk;
}

TFheGateBootstrappingSecretKeySet* new_random_gate_bootstrapping_secret_keyset(const TFheGateBootstrappingParameterSet* params) {
    oracle_params op = to_oracle_params(params);
    return wrap_keys(params, oracle_keygen(&op, S.seed_counter++), true);
}
/* a keyset around keys the test already holds (oracle_keys*, not owned): the GPU engine is loaded with the same arrays */
TFheGateBootstrappingSecretKeySet* refapi_keyset_from_oracle(const TFheGateBootstrappingParameterSet* params, void* keys) {
    return wrap_keys(params, (oracle_keys*)keys, false);
}
void delete_gate_bootstrapping_secret_keyset(TFheGateBootstrappingSecretKeySet* sk) {
    if (!sk) return;
    std::lock_guard<std::recursive_mutex> g(S.mu);
    if (sk->cloud.material) {
        if (S.keys == sk->cloud.material->keys) { flush_locked(); S.keys = nullptr; }
        if (sk->cloud.material->owned) oracle_keys_free(sk->cloud.material->keys);
        delete sk->cloud.material;
    }
    free(sk);
}
void delete_gate_bootstrapping_cloud_keyset(TFheGateBootstrappingCloudKeySet*) {}
void tfhe_random_generator_setSeed(uint32_t* values, int32_t size) {
    uint64_t s = 0x9e3779b97f4a7c15ULL;
    for (int32_t i = 0; i < size; ++i) s = (s ^ values[i]) * 0xbf58476d1ce4e5b9ULL;
    S.rng = s; S.seed_counter = s | 1;
}

/* ---- ciphertext arrays (main.cpp:59, Math.cpp:28,47) ---- */
LweSample* new_gate_bootstrapping_ciphertext_array(int32_t nbelems, const TFheGateBootstrappingParameterSet* params) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    const int32_t total = (nbelems > 0 ? nbelems : 0) + REFAPI_PAD;
    auto* arr = (LweSample*)calloc((size_t)total, sizeof(LweSample));
    for (int32_t i = 0; i < total; ++i) arr[i].peba1_n = params->n;
    S.arrays[arr] = total;
    return arr;
}
LweSample* new_gate_bootstrapping_ciphertext(const TFheGateBootstrappingParameterSet* params) {
    return new_gate_bootstrapping_ciphertext_array(1, params);
}
void delete_gate_bootstrapping_ciphertext_array(int32_t, LweSample* samples) {
    if (!samples) return;
    std::lock_guard<std::recursive_mutex> g(S.mu);
    auto it = S.arrays.find(samples);
    if (it == S.arrays.end()) die("delete of an array this library did not allocate");
    /* the values its samples named are swept by the next flush (nothing else names them after this) */
    S.arrays.erase(it);
    free(samples);
}
void delete_gate_bootstrapping_ciphertext(LweSample* s) { delete_gate_bootstrapping_ciphertext_array(1, s); }

/* ---- encryption (main.cpp:64,78) ---- */
void bootsSymEncrypt(LweSample* result, int32_t message, const TFheGateBootstrappingSecretKeySet* key) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    bind_keys(&key->cloud);
    int64_t h = new_value(0);
    val(h).words = alloc_words();
    oracle_encrypt_bit(S.keys, message ? 1 : 0, &S.rng, val(h).words);
    result->peba1_handle = (uint64_t)h;
}
int32_t bootsSymDecrypt(const LweSample* sample, const TFheGateBootstrappingSecretKeySet* key) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    bind_keys(&key->cloud);
    return oracle_decrypt_bit(S.keys, words_of(sample));
}

/* ---- gates (Math.cpp:31-43,58,84,153,158,165,170,233,260-261,277) ---- */
void bootsCONSTANT(LweSample* result, int32_t value, const TFheGateBootstrappingCloudKeySet* bk) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    bind_keys(bk);
    int64_t h = new_value(0);
    val(h).words = alloc_words();
    oracle_trivial(oracle_keys_params(S.keys), value ? 1 : 0, val(h).words);
    result->peba1_handle = (uint64_t)h;
    S.linear++;
}
void bootsNOT(LweSample* r, const LweSample* a, const TFheGateBootstrappingCloudKeySet* bk) { record(OPX_NOT, r, a, nullptr, nullptr, bk); }
void bootsCOPY(LweSample* r, const LweSample* a, const TFheGateBootstrappingCloudKeySet* bk) { record(OPX_COPY, r, a, nullptr, nullptr, bk); }
#define GATE2(NAME, OP) \
    void NAME(LweSample* r, const LweSample* a, const LweSample* b, const TFheGateBootstrappingCloudKeySet* bk) { record(OP, r, a, b, nullptr, bk); }
GATE2(bootsNAND, ORACLE_OP_NAND)
GATE2(bootsAND, ORACLE_OP_AND)
GATE2(bootsOR, ORACLE_OP_OR)
GATE2(bootsXOR, ORACLE_OP_XOR)
GATE2(bootsXNOR, ORACLE_OP_XNOR)
GATE2(bootsNOR, ORACLE_OP_NOR)
GATE2(bootsANDNY, ORACLE_OP_ANDNY)
GATE2(bootsANDYN, ORACLE_OP_ANDYN)
GATE2(bootsORNY, ORACLE_OP_ORNY)
GATE2(bootsORYN, ORACLE_OP_ORYN)
#undef GATE2
void bootsMUX(LweSample* r, const LweSample* a, const LweSample* b, const LweSample* c, const TFheGateBootstrappingCloudKeySet* bk) {
    record(ORACLE_OP_MUX, r, a, b, c, bk);
}

/* ---- control and raw access for the tests ---- */
void refapi_flush(void) { std::lock_guard<std::recursive_mutex> g(S.mu); flush_locked(); }
void refapi_set_threads(int threads) { S.threads = threads; }
void refapi_set_eager(int eager) { std::lock_guard<std::recursive_mutex> g(S.mu); flush_locked(); S.eager = eager; }
LweSample* refapi_sample_at(LweSample* array, int32_t index) { return array + index; }
void refapi_sample_set(LweSample* sample, const int32_t* words, const TFheGateBootstrappingCloudKeySet* bk) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    bind_keys(bk);
    int64_t h = new_value(0);
    val(h).words = alloc_words();
    memcpy(val(h).words, words, sizeof(int32_t) * (size_t)S.words);
    sample->peba1_handle = (uint64_t)h;
}
void refapi_sample_get(const LweSample* sample, const TFheGateBootstrappingCloudKeySet* bk, int32_t* words) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    bind_keys(bk);
    memcpy(words, words_of(sample), sizeof(int32_t) * (size_t)S.words);
}
typedef struct {
    uint64_t gates, pbs, linear, levels, flushes, max_level_width;
} refapi_stats;
void refapi_get_stats(refapi_stats* out) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    out->gates = S.gates; out->pbs = S.pbs; out->linear = S.linear; out->levels = S.levels;
    out->flushes = S.flushes; out->max_level_width = S.max_width;
}
void refapi_reset_stats(void) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    S.gates = S.pbs = S.linear = S.levels = S.flushes = S.max_width = 0;
}
/* widths of the levels of the last flush (level 0 = inputs, always 0); returns the count */
int64_t refapi_last_level_widths(uint64_t* out, int64_t cap) {
    std::lock_guard<std::recursive_mutex> g(S.mu);
    int64_t n = (int64_t)S.level_widths.size();
    for (int64_t i = 0; i < n && i < cap; ++i) out[i] = S.level_widths[(size_t)i];
    return n;
}

} // extern "C"
