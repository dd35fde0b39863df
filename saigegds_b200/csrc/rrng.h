// R's default random number stream, for the draws the reference takes through R closures:
//   set.seed(seed)            saige_fitnull.cpp:109-114, called at :631 and :676
//   rbinom(n, 1, 0.5)         :649 / :695   (Rademacher vectors of the Hutchinson trace estimator)
//   sample.int(n_var, n_var)  R/saige_main.r:509-511 (marker order of the variance-ratio step)
// under RNGkind("Mersenne-Twister", "Inversion", "Rounding") as fixed by the reference's tests
// (inst/unitTests/test_SAIGE.R:15).  Algorithm: MT19937 (Matsumoto & Nishimura) with R's seeding
// (R src/main/RNG.c: Randomize -> RNG_Init -> FixupSeeds) and R's [0,1) scaling + fixup.
// An R shim may replace this through sgb_set_callbacks() so that a user-chosen RNGkind is honoured.
#pragma once
#include <stdint.h>

#include <cmath>
#include <vector>

namespace sgb {

class RRng {
    static constexpr int kN = 624, kM = 397;
    uint32_t state_[kN];
    int pos_ = kN + 1;

    static uint32_t lcg(uint32_t s) { return 69069u * s + 1u; }

    void refill() {
        for (int k = 0; k < kN; k++) {
            uint32_t y = (state_[k] & 0x80000000u) | (state_[(k + 1) % kN] & 0x7fffffffu);
            uint32_t v = state_[(k + kM) % kN] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            state_[k] = v;
        }
        pos_ = 0;
    }

public:
    RRng() { set_seed(0); }

    void set_seed(uint32_t seed) {
        for (int j = 0; j < 50; j++) seed = lcg(seed);  // initial scrambling
        seed = lcg(seed);                               // i_seed[0] is the position slot ("dummy[0]")
        for (int j = 0; j < kN; j++) { seed = lcg(seed); state_[j] = seed; }
        pos_ = kN;                                      // FixupSeeds: mti = N -> regenerate on first draw
    }

    uint32_t next_u32() {
        if (pos_ >= kN) refill();
        uint32_t y = state_[pos_++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

    // unif_rand(): MT_genrand() scaled to [0,1), then fixup() into (0,1)
    double unif_rand() {
        const double eps = 2.328306437080797e-10;
        double v = next_u32() * 2.3283064365386963e-10;
        if (v <= 0.0) return 0.5 * eps;
        if (1.0 - v <= 0.0) return 1.0 - 0.5 * eps;
        return v;
    }

    // rbinom(1, 1, 0.5): the inversion branch of rbinom.c draws one uniform u and returns (u >= 0.5)
    int bernoulli_half() { return unif_rand() >= 0.5; }

    // sample.int(n, n) with sample.kind = "Rounding": partial Fisher-Yates, R_unif_index = floor(n * u)
    void sample_int(int32_t n, int32_t *out) {
        std::vector<int32_t> pool(n);
        for (int32_t i = 0; i < n; i++) pool[i] = i;
        int32_t left = n;
        for (int32_t i = 0; i < n; i++) {
            int32_t j = (int32_t)std::floor(left * unif_rand());
            out[i] = pool[j] + 1;
            pool[j] = pool[--left];
        }
    }
};

}  // namespace sgb
