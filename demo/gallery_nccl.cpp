// gallery_nccl.cpp — 1:N encrypted identification on N GPUs of one box, host side in C++ (no Python, no torch):
// the multi-GPU shape of the reference's protocol step (src/main.cpp:513-581: one encrypted probe, one enrolled
// template, Function_f) for a gallery of T templates.  BASELINE configs[3] / [4]; SURVEY.md §8e.
//
//   one process per GPU (this program forks them), each with its own libpeba1_tfhe / libpeba1_cuda context;
//   rank 0 = client + key holder: host keygen, encrypts probe, bound and gallery;
//   ncclBroadcast   of the DEVICE-format cloud keys (transformed bootstrapping key 62 MB + keyswitch key 62 MB) into
//                   peba1_cuda_key_device_buffers of the replicas — no key ever re-transformed, none bounced through hosts;
//   ncclBroadcast   of the probe and the bound ciphertexts, ncclSend/ncclRecv scatter of the gallery shards; what arrives
//                   is bound to ciphertext-pool rows by a device-side scatter (peba1_sample_set_list_device);
//   every rank records Function_f (include/peba1_math.h — the reference's own signature, std::vector<LweSample*>) for
//                   all of its templates into ONE batch: each of the 377 circuit levels is one launch pair per GPU;
//   ncclAllGather   of the encrypted match bits (one 2,524-byte ciphertext per template); rank 0 decrypts and checks.
// No collective sits inside the hot loop (templates are independent circuits): the timed region is recording + the
// batched levels + the all-gather.
//
//   make -C demo nccl && demo/peba1_gallery_nccl --gpus 2 --templates 16 [--mode 0|1|2|3] [--slots 128] [--seed 42]
#include <nccl.h>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "peba1_cuda.h"
#include "peba1_host.h"
#include "peba1_math.h"

#define CUDA_OK(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "rank %d: %s: %s\n", g_rank, #call, cudaGetErrorString(e_));           \
            _exit(2);                                                                              \
        }                                                                                          \
    } while (0)
#define NCCL_OK(call)                                                                              \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) {                                                                   \
            fprintf(stderr, "rank %d: %s: %s\n", g_rank, #call, ncclGetErrorString(r_));           \
            _exit(3);                                                                              \
        }                                                                                          \
    } while (0)
#define PEBA1_CHECK(call)                                                                             \
    do {                                                                                           \
        int r_ = (call);                                                                           \
        if (r_ != 0) {                                                                             \
            fprintf(stderr, "rank %d: %s -> %d: %s\n", g_rank, #call, r_, peba1_cuda_last_error()); \
            _exit(4);                                                                              \
        }                                                                                          \
    } while (0)

static int g_rank = 0;

struct Shared {  // anonymous shared mapping made before the fork: the NCCL id and the result line travel through it
    ncclUniqueId id;
    volatile int id_ready;
};

struct Options {
    int gpus = 1, templates = 8, mode = 0, slots = 128, bits = 8;
    uint64_t seed = 42;
};

static void barrier(ncclComm_t comm, cudaStream_t st, int* d_flag) {
    NCCL_OK(ncclAllReduce(d_flag, d_flag, 1, ncclInt, ncclMax, comm, st));
    CUDA_OK(cudaStreamSynchronize(st));
}

static int run_rank(const Options& o, int rank, Shared* sh) {
    g_rank = rank;
    const int world = o.gpus, slots = o.slots, bits = o.bits;
    CUDA_OK(cudaSetDevice(rank));
    peba1_host_set_device(rank);
    peba1_set_mode(o.mode);
    if (rank == 0) {
        NCCL_OK(ncclGetUniqueId(&sh->id));
        __sync_synchronize();
        sh->id_ready = 1;
    } else {
        while (!sh->id_ready) usleep(1000);
        __sync_synchronize();
    }
    ncclComm_t comm;
    NCCL_OK(ncclCommInitRank(&comm, world, sh->id, rank));
    cudaStream_t st;
    CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int* d_flag = nullptr;
    CUDA_OK(cudaMalloc(&d_flag, sizeof(int)));
    CUDA_OK(cudaMemset(d_flag, 0, sizeof(int)));

    // ---- keys (main.cpp:21-23 on rank 0; replicas hold no key material on the host)
    TFheGateBootstrappingParameterSet* params = new_default_gate_bootstrapping_parameters(128);
    TFheGateBootstrappingSecretKeySet* key =
        rank == 0 ? peba1_new_secret_keyset_seeded(params, o.seed) : peba1_new_placeholder_keyset(params);
    const TFheGateBootstrappingCloudKeySet* cloud = &key->cloud;
    peba1_cuda_ctx* ctx = (peba1_cuda_ctx*)peba1_cuda_context_of(cloud);
    if (!ctx) { fprintf(stderr, "rank %d: no CUDA context (this library has no CPU path)\n", rank); _exit(5); }
    PEBA1_CHECK(peba1_cuda_set_stream(ctx, st));  // gate launches, NCCL and the copies below share one stream: ordered by it
    void *d_bk = nullptr, *d_ksk = nullptr;
    uint64_t bk_bytes = 0, ksk_bytes = 0;
    PEBA1_CHECK(peba1_cuda_key_device_buffers(ctx, &d_bk, &bk_bytes, &d_ksk, &ksk_bytes));
    const auto tk0 = std::chrono::steady_clock::now();
    NCCL_OK(ncclBroadcast(d_bk, d_bk, bk_bytes, ncclChar, 0, comm, st));
    NCCL_OK(ncclBroadcast(d_ksk, d_ksk, ksk_bytes, ncclChar, 0, comm, st));
    CUDA_OK(cudaStreamSynchronize(st));
    const double key_bcast_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - tk0).count();
    if (rank != 0) PEBA1_CHECK(peba1_cuda_keys_mark_loaded(ctx));

    // ---- the client's data: probe, bound, gallery (a quarter of the templates are near-copies of the probe)
    const int words = params->n + 1;
    const int per = (o.templates + world - 1) / world, total = per * world;
    const size_t vec_words = (size_t)slots * bits * words;
    const uint64_t bound_value = 512;
    int32_t *d_probe = nullptr, *d_bound = nullptr, *d_mine = nullptr, *d_all = nullptr;
    CUDA_OK(cudaMalloc(&d_probe, vec_words * sizeof(int32_t)));
    CUDA_OK(cudaMalloc(&d_bound, (size_t)24 * words * sizeof(int32_t)));
    CUDA_OK(cudaMalloc(&d_mine, (size_t)per * vec_words * sizeof(int32_t)));
    std::vector<int> expect((size_t)total);
    if (rank == 0) {
        std::mt19937_64 rng(o.seed * 7919 + 1);
        std::vector<int> probe((size_t)slots);
        for (int& v : probe) v = 1 + (int)(rng() % 255);  // no zero byte: the reference's subtraction defect (quirk Q1) stays silent
        std::vector<int32_t> pbits((size_t)slots * bits), bbits(24), gbits((size_t)total * slots * bits);
        for (int i = 0; i < slots; ++i)
            for (int b = 0; b < bits; ++b) pbits[(size_t)i * bits + b] = (probe[i] >> b) & 1;
        for (int b = 0; b < 24; ++b) bbits[b] = (int32_t)((bound_value >> b) & 1);
        for (int t = 0; t < total; ++t) {
            uint64_t d2 = 0;
            for (int i = 0; i < slots; ++i) {
                int v = (int)(rng() % 256);
                if (t % 4 == 1) { v = probe[i] + (int)(rng() % 5) - 2; v = v < 0 ? 0 : v > 255 ? 255 : v; }
                d2 += (uint64_t)((v - probe[i]) * (v - probe[i]));
                for (int b = 0; b < bits; ++b) gbits[((size_t)t * slots + i) * bits + b] = (v >> b) & 1;
            }
            expect[t] = d2 > bound_value ? 1 : 0;  // polarity of the reference's comparator (quirk Q3): 1 = NO match
        }
        std::vector<int32_t> w(std::max((size_t)total * vec_words, vec_words));
        peba1_encrypt_bits(key, pbits.data(), (int64_t)pbits.size(), 501, w.data());
        CUDA_OK(cudaMemcpy(d_probe, w.data(), vec_words * sizeof(int32_t), cudaMemcpyHostToDevice));
        peba1_encrypt_bits(key, bbits.data(), 24, 502, w.data());
        CUDA_OK(cudaMemcpy(d_bound, w.data(), (size_t)24 * words * sizeof(int32_t), cudaMemcpyHostToDevice));
        peba1_encrypt_bits(key, gbits.data(), (int64_t)gbits.size(), 503, w.data());
        CUDA_OK(cudaMalloc(&d_all, (size_t)total * vec_words * sizeof(int32_t)));
        CUDA_OK(cudaMemcpy(d_all, w.data(), (size_t)total * vec_words * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    NCCL_OK(ncclBroadcast(d_probe, d_probe, vec_words, ncclInt32, 0, comm, st));
    NCCL_OK(ncclBroadcast(d_bound, d_bound, (size_t)24 * words, ncclInt32, 0, comm, st));
    NCCL_OK(ncclGroupStart());  // scatter of the shards
    if (rank == 0)
        for (int r = 0; r < world; ++r)
            NCCL_OK(ncclSend(d_all + (size_t)r * per * vec_words, (size_t)per * vec_words, ncclInt32, r, comm, st));
    NCCL_OK(ncclRecv(d_mine, (size_t)per * vec_words, ncclInt32, 0, comm, st));
    NCCL_OK(ncclGroupEnd());
    CUDA_OK(cudaStreamSynchronize(st));

    // ---- ciphertext arrays as the reference's caller holds them (main.cpp:59), bound to what arrived on the device
    auto make_vector = [&](std::vector<LweSample*>& v, int n, int len) {
        v.resize((size_t)n);
        for (auto& a : v) a = new_gate_bootstrapping_ciphertext_array(len, params);
    };
    std::vector<LweSample*> probe_v, bound_v;
    make_vector(probe_v, slots, bits);
    peba1_sample_set_list_device(probe_v.data(), slots, bits, d_probe, cloud);
    make_vector(bound_v, 1, 24);
    peba1_sample_set_list_device(bound_v.data(), 1, 24, d_bound, cloud);
    std::vector<std::vector<LweSample*>> templ((size_t)per);
    std::vector<LweSample*> flat;
    for (int t = 0; t < per; ++t) {
        make_vector(templ[t], slots, bits);
        flat.insert(flat.end(), templ[t].begin(), templ[t].end());
    }
    peba1_sample_set_list_device(flat.data(), (int32_t)flat.size(), bits, d_mine, cloud);
    std::vector<LweSample*> result((size_t)per);
    for (auto& r : result) r = new_gate_bootstrapping_ciphertext_array(24, params);

    // ---- the timed region: Function_f for every template of the shard in one batch, then the all-gather
    peba1_reset_stats();
    int32_t *d_send = nullptr, *d_recv = nullptr;
    const size_t row = (size_t)words;
    CUDA_OK(cudaMalloc(&d_send, (size_t)per * row * sizeof(int32_t)));
    CUDA_OK(cudaMalloc(&d_recv, (size_t)total * row * sizeof(int32_t)));
    barrier(comm, st, d_flag);
    const auto t0 = std::chrono::steady_clock::now();
    peba1_batch_begin();
    for (int t = 0; t < per; ++t) Function_f(result[t], probe_v, templ[t], bound_v[0], bits, cloud);  // Math.cpp:379
    peba1_batch_end();
    void* pool = nullptr;
    int64_t stride = 0, cap = 0;
    for (int t = 0; t < per; ++t) {  // result_b[0] of every template: pool row -> send buffer, device to device
        const int64_t slot = peba1_sample_device_slot(&result[t][0], cloud);
        if (slot < 0) { fprintf(stderr, "rank %d: result %d is not pool-resident\n", rank, t); _exit(6); }
        PEBA1_CHECK(peba1_cuda_pool_info(ctx, &pool, &stride, &cap));
        CUDA_OK(cudaMemcpyAsync(d_send + (size_t)t * row, (const int32_t*)pool + slot * stride, row * sizeof(int32_t),
                                cudaMemcpyDeviceToDevice, st));
    }
    NCCL_OK(ncclAllGather(d_send, d_recv, (size_t)per * row, ncclInt32, comm, st));
    CUDA_OK(cudaStreamSynchronize(st));
    barrier(comm, st, d_flag);
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    peba1_host_stats hs;
    peba1_get_stats(&hs);
    int rc = 0;
    if (rank == 0) {
        std::vector<int32_t> all((size_t)total * row), got((size_t)total);
        CUDA_OK(cudaMemcpy(all.data(), d_recv, all.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        peba1_decrypt_words(key, all.data(), total, got.data());
        int wrong = 0, matches = 0;
        for (int t = 0; t < total; ++t) { wrong += got[t] != expect[t]; matches += got[t] == 0; }
        printf("{\"program\": \"demo/gallery_nccl.cpp\", \"n_gpus\": %d, \"templates\": %d, \"templates_per_gpu\": %d, "
               "\"circuit_mode\": %d, \"seconds\": %.4f, \"face_matches_per_s\": %.4f, \"pbs_per_gpu\": %llu, "
               "\"pbs_per_s\": %.1f, \"levels\": %llu, \"key_broadcast_seconds\": %.4f, \"key_broadcast_bytes\": %llu, "
               "\"match_bits_gathered_bytes\": %zu, \"matches_found\": %d, \"verified\": %s}\n",
               world, total, per, o.mode, seconds, total / seconds, (unsigned long long)hs.executed_pbs,
               (double)hs.executed_pbs * world / seconds, (unsigned long long)hs.levels, key_bcast_s,
               (unsigned long long)(bk_bytes + ksk_bytes), (size_t)total * row * sizeof(int32_t), matches,
               wrong == 0 ? "true" : "false");
        fflush(stdout);
        rc = wrong == 0 ? 0 : 1;
    }
    for (auto& a : probe_v) delete_gate_bootstrapping_ciphertext_array(bits, a);
    for (auto& a : flat) delete_gate_bootstrapping_ciphertext_array(bits, a);
    delete_gate_bootstrapping_ciphertext_array(24, bound_v[0]);
    for (auto& r : result) delete_gate_bootstrapping_ciphertext_array(24, r);
    cudaFree(d_probe); cudaFree(d_bound); cudaFree(d_mine); cudaFree(d_all); cudaFree(d_send); cudaFree(d_recv); cudaFree(d_flag);
    NCCL_OK(ncclCommDestroy(comm));
    return rc;
}

int main(int argc, char** argv) {
    Options o;
    for (int i = 1; i + 1 < argc; i += 2) {
        const long long v = atoll(argv[i + 1]);
        if (!strcmp(argv[i], "--gpus")) o.gpus = (int)v;
        else if (!strcmp(argv[i], "--templates")) o.templates = (int)v;
        else if (!strcmp(argv[i], "--mode")) o.mode = (int)v;
        else if (!strcmp(argv[i], "--slots")) o.slots = (int)v;
        else if (!strcmp(argv[i], "--seed")) o.seed = (uint64_t)v;
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
    }
    if (o.gpus < 1 || o.templates < 1 || o.slots < 1 || o.mode < 0 || o.mode > 3) { fprintf(stderr, "bad arguments\n"); return 2; }
    Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh == MAP_FAILED) { perror("mmap"); return 2; }
    memset(sh, 0, sizeof *sh);
    // one process per GPU, forked BEFORE anything touches CUDA
    std::vector<pid_t> kids;
    for (int r = 1; r < o.gpus; ++r) {
        const pid_t pid = fork();
        if (pid < 0) { perror("fork"); return 2; }
        if (pid == 0) _exit(run_rank(o, r, sh));
        kids.push_back(pid);
    }
    int rc = run_rank(o, 0, sh);
    for (pid_t pid : kids) {
        int status = 0;
        waitpid(pid, &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = rc ? rc : 7;
    }
    return rc;
}
