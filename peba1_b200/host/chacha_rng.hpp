// chacha_rng.hpp — the host side's random generator: ChaCha20 (RFC 8439 block function) in counter mode, keyed with
// 256 bits.  Key bits, LWE masks and noise all come from it.  The process-wide generator is keyed from the OS entropy
// pool (getrandom) at first use; upstream TFHE's default engine is unseeded and the reference never seeds it
// (SURVEY.md §8b), so a drop-in that kept that behaviour would hand every process the same secret key.
// Deterministic streams exist only where the caller names a seed (peba1_new_secret_keyset_seeded, peba1_encrypt_bits,
// tfhe_random_generator_setSeed): tests and benchmarks.
#pragma once

#include <stdint.h>
#include <string.h>
#include <sys/random.h>

#include <limits>

namespace peba1 {

class ChaChaRng {
public:
    using result_type = uint64_t;
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return std::numeric_limits<uint64_t>::max(); }

    ChaChaRng() { memset(key_, 0, sizeof key_); }
    // independent streams under one key: `stream` is the 64-bit nonce
    ChaChaRng(const uint32_t key[8], uint64_t stream) : stream_(stream) { memcpy(key_, key, sizeof key_); }
    // a 64-bit seed stretched to a key (deterministic test/bench streams; NOT 128-bit secure, by construction)
    static ChaChaRng from_seed(uint64_t seed, uint64_t stream = 0) {
        uint32_t key[8];
        expand_seed(seed, key);
        return ChaChaRng(key, stream);
    }
    static void expand_seed(uint64_t seed, uint32_t key[8]) {
        for (int i = 0; i < 4; ++i) {  // splitmix64
            uint64_t z = (seed += 0x9E3779B97F4A7C15ULL);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            z ^= z >> 31;
            key[2 * i] = (uint32_t)z;
            key[2 * i + 1] = (uint32_t)(z >> 32);
        }
    }
    // 256 bits from the OS; returns false when the pool cannot be read (the caller aborts: no silent weak keys)
    static bool os_key(uint32_t key[8]) {
        size_t got = 0;
        while (got < 32) {
            const ssize_t r = getrandom((char*)key + got, 32 - got, 0);
            if (r <= 0) return false;
            got += (size_t)r;
        }
        return true;
    }
    const uint32_t* key() const { return key_; }
    void seek(uint64_t block) { counter_ = block; pos_ = 8; }  // known-answer tests (RFC 8439 §2.3.2)

    uint64_t operator()() {
        if (pos_ >= 8) refill();
        return buf_[pos_++];
    }

private:
    static inline uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    static inline void qr(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
        a += b; d ^= a; d = rotl(d, 16);
        c += d; b ^= c; b = rotl(b, 12);
        a += b; d ^= a; d = rotl(d, 8);
        c += d; b ^= c; b = rotl(b, 7);
    }
    void refill() {
        uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        memcpy(s + 4, key_, sizeof key_);
        s[12] = (uint32_t)counter_;
        s[13] = (uint32_t)(counter_ >> 32);
        s[14] = (uint32_t)stream_;
        s[15] = (uint32_t)(stream_ >> 32);
        uint32_t x[16];
        memcpy(x, s, sizeof x);
        for (int r = 0; r < 10; ++r) {
            qr(x[0], x[4], x[8], x[12]); qr(x[1], x[5], x[9], x[13]); qr(x[2], x[6], x[10], x[14]); qr(x[3], x[7], x[11], x[15]);
            qr(x[0], x[5], x[10], x[15]); qr(x[1], x[6], x[11], x[12]); qr(x[2], x[7], x[8], x[13]); qr(x[3], x[4], x[9], x[14]);
        }
        for (int i = 0; i < 8; ++i) buf_[i] = (uint64_t)(x[2 * i] + s[2 * i]) | ((uint64_t)(x[2 * i + 1] + s[2 * i + 1]) << 32);
        ++counter_;
        pos_ = 0;
    }
    uint32_t key_[8];
    uint64_t stream_ = 0, counter_ = 0;
    uint64_t buf_[8];
    int pos_ = 8;
};

}  // namespace peba1
