// twiddles_host.h — host-side construction of the per-thread twiddle tables (long double precision).
#pragma once
#include <math.h>
#include "blind_rotate.cuh"

namespace peba1 {

inline void build_twiddle_tables(TwiddleTables& T) {
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int tid = 0; tid < 64; ++tid)
        for (int j = 0; j < 8; ++j) {
            const int e = (tid * (1 + 4 * j)) % 2048;  // zeta^e, zeta = exp(i pi / 1024)
            T.tw1[tid][j] = make_double2((double)cosl(PI * e / 1024.0L), (double)sinl(PI * e / 1024.0L));
        }
    for (int c = 0; c < 8; ++c)
        for (int j = 0; j < 8; ++j) {
            const int e = (c * j) % 64;  // w64^e
            T.tw2[c][j] = make_double2((double)cosl(2 * PI * e / 64.0L), (double)sinl(2 * PI * e / 64.0L));
        }
}

inline void load_team_twiddles(const TwiddleTables& T, int tid, TeamTwiddles& tw) {
    for (int j = 0; j < 8; ++j) { tw.tw1[j] = T.tw1[tid][j]; tw.tw2[j] = T.tw2[tid & 7][j]; }
}

}  // namespace peba1
