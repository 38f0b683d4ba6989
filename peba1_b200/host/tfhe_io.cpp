// tfhe_io.cpp — <tfhe/tfhe_io.h>: key and ciphertext serialisation under upstream TFHE's function names
// (formats and their parity status are described in the header).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "peba1_host.h"
#include "tfhe/tfhe_gate_bootstrapping_functions.h"
#include "tfhe/tfhe_io.h"

namespace {

[[noreturn]] void io_die(const char* what) {
    fprintf(stderr, "peba1 tfhe_io: %s\n", what);
    abort();
}
void put(FILE* F, const void* p, size_t bytes) {
    if (!F) io_die("null FILE");
    if (bytes && fwrite(p, 1, bytes, F) != bytes) io_die("write failed");
}
void get(FILE* F, void* p, size_t bytes) {
    if (!F) io_die("null FILE");
    if (bytes && fread(p, 1, bytes, F) != bytes) io_die("unexpected end of file");
}
uint64_t fnv1a(uint64_t h, const void* p, size_t bytes) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < bytes; ++i) h = (h ^ b[i]) * 0x100000001B3ULL;
    return h;
}
constexpr uint64_t kFnvInit = 0xCBF29CE484222325ULL;

// container: "PEBA1IO\0", u32 version, u32 kind, the fixed-size parameter record, payload arrays, u64 FNV-1a of
// everything after the magic
constexpr char kMagic[8] = {'P', 'E', 'B', 'A', '1', 'I', 'O', 0};
constexpr uint32_t kVersion = 1;
enum Kind : uint32_t { kParams = 1, kCloud = 2, kSecret = 3 };
constexpr int32_t kLweSampleTypeUid = 42;  // upstream's LWE_SAMPLE_TYPE_UID [UPSTREAM-RECALL]

struct ParamRecord {  // explicit, padding-free layout (the in-memory struct's padding is not written)
    int32_t ks_t, ks_basebit, n, N, k, l, Bgbit, minimum_lambda;
    double ks_stdev, bk_stdev, max_stdev;
};
static_assert(sizeof(ParamRecord) == 56, "parameter record layout");

ParamRecord to_record(const TFheGateBootstrappingParameterSet& p) {
    return ParamRecord{p.ks_t, p.ks_basebit, p.n, p.N, p.k, p.l, p.Bgbit, p.minimum_lambda, p.ks_stdev, p.bk_stdev, p.max_stdev};
}
void from_record(const ParamRecord& r, TFheGateBootstrappingParameterSet& p) {
    p.ks_t = r.ks_t; p.ks_basebit = r.ks_basebit; p.n = r.n; p.N = r.N; p.k = r.k; p.l = r.l; p.Bgbit = r.Bgbit;
    p.minimum_lambda = r.minimum_lambda; p.ks_stdev = r.ks_stdev; p.bk_stdev = r.bk_stdev; p.max_stdev = r.max_stdev;
}
bool sane(const ParamRecord& r) {
    return r.n > 0 && r.n < 4096 && r.N > 0 && r.N <= 4096 && r.k == 1 && r.l > 0 && r.l <= 8 && r.Bgbit > 0 &&
           r.Bgbit * r.l <= 32 && r.ks_t > 0 && r.ks_basebit > 0 && r.ks_t * r.ks_basebit <= 31;
}
size_t bk_words(const ParamRecord& r) { return (size_t)r.n * (r.k + 1) * r.l * 2 * r.N; }
size_t ksk_words(const ParamRecord& r) { return (size_t)r.N * r.ks_t * ((size_t)1 << r.ks_basebit) * (r.n + 1); }

struct Writer {
    FILE* F;
    uint64_t h = kFnvInit;
    void w(const void* p, size_t bytes) { put(F, p, bytes); h = fnv1a(h, p, bytes); }
    void header(uint32_t kind, const ParamRecord& r) {
        put(F, kMagic, sizeof kMagic);
        w(&kVersion, 4);
        w(&kind, 4);
        w(&r, sizeof r);
    }
    void finish() { put(F, &h, 8); }
};
struct Reader {
    FILE* F;
    uint64_t h = kFnvInit;
    void r(void* p, size_t bytes) { get(F, p, bytes); h = fnv1a(h, p, bytes); }
    ParamRecord header(uint32_t want_kind) {
        char magic[8];
        get(F, magic, sizeof magic);
        if (memcmp(magic, kMagic, sizeof magic)) io_die("not a peba1 key/parameter file (bad magic)");
        uint32_t version, kind;
        r(&version, 4);
        r(&kind, 4);
        if (version != kVersion) io_die("unsupported file version");
        if (kind != want_kind) io_die("file holds a different kind of object");
        ParamRecord rec;
        r(&rec, sizeof rec);
        if (!sane(rec)) io_die("implausible parameter set in file");
        return rec;
    }
    void finish() {
        uint64_t stored;
        get(F, &stored, 8);
        if (stored != h) io_die("checksum mismatch (corrupt or truncated file)");
    }
};

}  // namespace

extern "C" {

void export_tfheGateBootstrappingParameterSet_toFile(FILE* F, const TFheGateBootstrappingParameterSet* params) {
    if (!params) io_die("null parameter set");
    Writer W{F};
    W.header(kParams, to_record(*params));
    W.finish();
}
TFheGateBootstrappingParameterSet* new_tfheGateBootstrappingParameterSet_fromFile(FILE* F) {
    Reader R{F};
    const ParamRecord rec = R.header(kParams);
    R.finish();
    TFheGateBootstrappingParameterSet* p = new_default_gate_bootstrapping_parameters(rec.minimum_lambda > 128 ? 128 : rec.minimum_lambda);
    from_record(rec, *p);
    return p;
}

void export_tfheGateBootstrappingCloudKeySet_toFile(FILE* F, const TFheGateBootstrappingCloudKeySet* cloud) {
    if (!cloud || !cloud->params) io_die("null cloud keyset");
    const int32_t *bk = peba1_keyset_bk(cloud), *ksk = peba1_keyset_ksk(cloud);
    if (!bk || !ksk) io_die("this cloud keyset holds no key material (placeholder, or its secret keyset was deleted)");
    const ParamRecord rec = to_record(*cloud->params);
    Writer W{F};
    W.header(kCloud, rec);
    W.w(bk, bk_words(rec) * sizeof(int32_t));
    W.w(ksk, ksk_words(rec) * sizeof(int32_t));
    W.finish();
}
TFheGateBootstrappingCloudKeySet* new_tfheGateBootstrappingCloudKeySet_fromFile(FILE* F) {
    Reader R{F};
    const ParamRecord rec = R.header(kCloud);
    std::vector<int32_t> bk(bk_words(rec)), ksk(ksk_words(rec));
    R.r(bk.data(), bk.size() * sizeof(int32_t));
    R.r(ksk.data(), ksk.size() * sizeof(int32_t));
    R.finish();
    TFheGateBootstrappingParameterSet p{};
    from_record(rec, p);
    return peba1_new_cloud_keyset_from_arrays(&p, bk.data(), ksk.data());
}

void export_tfheGateBootstrappingSecretKeySet_toFile(FILE* F, const TFheGateBootstrappingSecretKeySet* keyset) {
    if (!keyset || !keyset->params || !keyset->lwe_key || !keyset->tlwe_key) io_die("null or key-less secret keyset");
    const int32_t *bk = peba1_keyset_bk(&keyset->cloud), *ksk = peba1_keyset_ksk(&keyset->cloud);
    if (!bk || !ksk) io_die("this keyset holds no cloud key material");
    const ParamRecord rec = to_record(*keyset->params);
    Writer W{F};
    W.header(kSecret, rec);
    W.w(keyset->lwe_key, (size_t)rec.n * sizeof(int32_t));
    W.w(keyset->tlwe_key, (size_t)rec.N * sizeof(int32_t));
    W.w(bk, bk_words(rec) * sizeof(int32_t));
    W.w(ksk, ksk_words(rec) * sizeof(int32_t));
    W.finish();
}
TFheGateBootstrappingSecretKeySet* new_tfheGateBootstrappingSecretKeySet_fromFile(FILE* F) {
    Reader R{F};
    const ParamRecord rec = R.header(kSecret);
    std::vector<int32_t> lwe((size_t)rec.n), tlwe((size_t)rec.N), bk(bk_words(rec)), ksk(ksk_words(rec));
    R.r(lwe.data(), lwe.size() * sizeof(int32_t));
    R.r(tlwe.data(), tlwe.size() * sizeof(int32_t));
    R.r(bk.data(), bk.size() * sizeof(int32_t));
    R.r(ksk.data(), ksk.size() * sizeof(int32_t));
    R.finish();
    for (int32_t v : lwe) if (v != 0 && v != 1) io_die("LWE key is not binary");
    for (int32_t v : tlwe) if (v != 0 && v != 1) io_die("ring key is not binary");
    TFheGateBootstrappingParameterSet p{};
    from_record(rec, p);
    return peba1_new_keyset_from_arrays(&p, lwe.data(), tlwe.data(), bk.data(), ksk.data());
}

void export_gate_bootstrapping_ciphertext_toFile(FILE* F, const LweSample* sample,
                                                 const TFheGateBootstrappingParameterSet* params) {
    if (!sample || !params) io_die("null argument");
    if (sample->peba1_n != params->n) io_die("ciphertext and parameter set disagree on n");
    std::vector<int32_t> w((size_t)params->n + 1);
    peba1_sample_get_words(sample, w.data());  // executes pending gates if the value is not materialised yet
    put(F, &kLweSampleTypeUid, sizeof(int32_t));
    put(F, w.data(), sizeof(int32_t) * (size_t)params->n);
    put(F, &w[(size_t)params->n], sizeof(int32_t));
    const double variance = sample->current_variance;
    put(F, &variance, sizeof(double));
}
void import_gate_bootstrapping_ciphertext_fromFile(FILE* F, LweSample* sample,
                                                   const TFheGateBootstrappingParameterSet* params) {
    if (!sample || !params) io_die("null argument");
    if (sample->peba1_n != params->n) io_die("ciphertext and parameter set disagree on n");
    int32_t uid;
    get(F, &uid, sizeof uid);
    if (uid != kLweSampleTypeUid) io_die("not an LWE sample (bad type id)");
    std::vector<int32_t> w((size_t)params->n + 1);
    get(F, w.data(), sizeof(int32_t) * ((size_t)params->n + 1));
    double variance;
    get(F, &variance, sizeof variance);
    peba1_sample_set(sample, w.data());
    sample->current_variance = variance;
}

}  // extern "C"
