// ORACLE (test infrastructure).  Samplers: ZSobol (bit-exact restatement) and the counter RNG standing in for ThreadRng.
// Follows /root/reference/renderer/src/sampler/{z_sobol_sampler,random_sampler}.rs.
// Pinned by the known-answer vectors in SURVEY.md Appendix B (tests/test_sobol.py) and the in-tree Sobol matrices.
#pragma once
#include <cstdint>

#include "omath.h"

namespace orc {

struct Tables;  // ospectrum.h

// z_sobol_sampler.rs:3-29
struct FastOwenScrambler {
    uint32_t seed;
    static uint32_t reverse_bits_32(uint32_t n) {
        n = (n >> 16) | (n << 16);
        n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
        n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
        n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
        n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
        return n;
    }
    uint32_t randomize(uint32_t v) const {
        v = reverse_bits_32(v);
        v ^= v * 0x3d20adeau;
        v += seed;
        v *= (seed >> 16) | 1u;
        v ^= v * 0x05526c56u;
        v ^= v * 0x53a22864u;
        return reverse_bits_32(v);
    }
};

struct SamplerBase {
    virtual ~SamplerBase() {}
    virtual void start_pixel_sample(uint32_t px, uint32_t py, uint32_t sample_index) = 0;
    virtual float get_1d() = 0;
    virtual Vec2 get_2d() = 0;
    Vec2 get_2d_pixel() { return get_2d(); }
    // key of the independent streams used where the reference calls rand::rng() inside shading (generalized_schlick.rs:901);
    // the reference's ThreadRng is OS-seeded, so any fixed stream is an equally valid stand-in.  One stream per
    // (path, bounce, call site), defined identically in the device code (csrc/dcommon.cuh aux_rng).
    virtual uint32_t aux_base() const = 0;
};

// z_sobol_sampler.rs:33-235
struct ZSobolSampler : SamplerBase {
    const uint32_t* matrices;  // 2 x 52 words (sobol_matrices.rs:7, dims 0 and 1)
    uint32_t dimension = 0, seed = 0, log2_spp = 0, n_base4_digits = 0, morton_index = 0;
    uint32_t aux_key = 0;

    static uint32_t log2_int(uint32_t v) { return v == 0 ? 0 : 31 - (uint32_t)__builtin_clz(v); }
    static uint32_t round_up_pow2(uint32_t v) { return v <= 1 ? 1u : 1u << (32 - __builtin_clz(v - 1)); }
    static uint64_t left_shift2(uint64_t x) {
        x &= 0xffffffffull;
        x = (x ^ (x << 16)) & 0x0000ffff0000ffffull;
        x = (x ^ (x << 8)) & 0x00ff00ff00ff00ffull;
        x = (x ^ (x << 4)) & 0x0f0f0f0f0f0f0f0full;
        x = (x ^ (x << 2)) & 0x3333333333333333ull;
        x = (x ^ (x << 1)) & 0x5555555555555555ull;
        return x;
    }
    static uint32_t encode_morton2(uint32_t x, uint32_t y) { return ((uint32_t)left_shift2(y) << 1) | (uint32_t)left_shift2(x); }
    static uint64_t mix_bits(uint64_t v) {
        v ^= v >> 31;
        v *= 0x7fb5d329728ea185ull;
        v ^= v >> 27;
        v *= 0x81dadef4bc2dd44dull;
        v ^= v >> 33;
        return v;
    }
    // MurmurHash64A of the 8 bytes (dimension LE, seed LE) (z_sobol_sampler.rs:76-99)
    static uint64_t hash(uint32_t dimension, uint32_t seed) {
        const uint64_t M = 0xc6a4a7935bd1e995ull;
        const int R = 47;
        uint64_t h = 8ull * M;
        uint64_t k = (uint64_t)dimension | ((uint64_t)seed << 32);
        k *= M;
        k ^= k >> R;
        k *= M;
        h ^= k;
        h *= M;
        h ^= h >> R;
        h *= M;
        h ^= h >> R;
        return h;
    }

    ZSobolSampler(const uint32_t* mats, uint32_t spp, uint32_t w, uint32_t h, uint32_t seed_) : matrices(mats), seed(seed_) {
        log2_spp = log2_int(spp);
        uint32_t res = round_up_pow2(w > h ? w : h);
        uint32_t log4_spp = (log2_spp + 1) / 2;
        n_base4_digits = log2_int(res) + log4_spp;
    }

    void start_pixel_sample(uint32_t px, uint32_t py, uint32_t sample_index) override {
        dimension = 0;
        // u32 shift silently drops high bits; Rust `<<` on u32 with shift < 32 wraps the value bits out
        morton_index = (encode_morton2(px, py) << log2_spp) | sample_index;
        aux_key = (px * 0x9e3779b9u) ^ (py * 0x85ebca6bu) ^ (sample_index * 0xc2b2ae35u) ^ seed ^ 0x5bd1e995u;
    }

    uint64_t get_sample_index() const {
        static const uint8_t PERM[24][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 1, 3}, {0, 2, 3, 1}, {0, 3, 2, 1}, {0, 3, 1, 2},
                                            {1, 0, 2, 3}, {1, 0, 3, 2}, {1, 2, 0, 3}, {1, 2, 3, 0}, {1, 3, 2, 0}, {1, 3, 0, 2},
                                            {2, 1, 0, 3}, {2, 1, 3, 0}, {2, 0, 1, 3}, {2, 0, 3, 1}, {2, 3, 0, 1}, {2, 3, 1, 0},
                                            {3, 1, 2, 0}, {3, 1, 0, 2}, {3, 2, 1, 0}, {3, 2, 0, 1}, {3, 0, 2, 1}, {3, 0, 1, 2}};
        uint64_t sample_index = 0;
        bool pow2_samples = (log2_spp & 1) == 1;
        int last_digit = pow2_samples ? 1 : 0;
        int i = (int)n_base4_digits - 1;
        while (i >= last_digit) {
            int digit_shift = 2 * i - (pow2_samples ? 1 : 0);
            uint64_t digit = ((uint64_t)morton_index >> digit_shift) & 3;
            uint64_t higher_digits = (uint64_t)morton_index >> (digit_shift + 2);
            uint64_t p = (mix_bits(higher_digits ^ (0x55555555ull * (uint64_t)dimension)) >> 24) % 24;
            digit = PERM[p][digit];
            sample_index |= digit << digit_shift;
            i -= 1;
        }
        if (pow2_samples) {
            // quirk: `& i as u64` with i == 0 after the loop (pbrt has `& 1`)
            uint64_t digit = (uint64_t)morton_index & (uint64_t)(int64_t)i;
            sample_index |= digit ^ ((mix_bits(((uint64_t)morton_index >> 1) ^ (0x55555555ull * (uint64_t)dimension))) & 1);
        }
        return sample_index;
    }

    float sobol_sample(uint64_t a, int dim, uint32_t scramble_seed) const {
        uint32_t v = 0;
        int i = dim * 52;
        while (a != 0) {
            if (a & 1) v ^= matrices[i];
            a >>= 1;
            i += 1;
        }
        v = FastOwenScrambler{scramble_seed}.randomize(v);
        float f = (float)v * 2.3283064365386963e-10f;  // 0x1p-32
        const float ONE_MINUS_EPS = 0.99999994f;        // 0x3f7fffff
        return f < ONE_MINUS_EPS ? f : ONE_MINUS_EPS;
    }

    float get_1d() override {
        uint64_t idx = get_sample_index();
        dimension += 1;
        uint64_t h = hash(dimension, seed);
        return sobol_sample(idx, 0, (uint32_t)h);
    }
    Vec2 get_2d() override {
        uint64_t idx = get_sample_index();
        dimension += 2;
        uint64_t bits = hash(dimension, seed);
        Vec2 r;
        r.x = sobol_sample(idx, 0, (uint32_t)bits);
        r.y = sobol_sample(idx, 1, (uint32_t)(bits >> 32));
        return r;
    }
    uint32_t aux_base() const override { return aux_key; }
};

// Counter RNG shared by definition with the device code ("pcg4d"-style hash of (key, counter)); stands in for
// rand::ThreadRng (random_sampler.rs:5-43, generalized_schlick.rs:901), which is OS-seeded and not reproducible:
// only statistical parity with the reference is possible for these streams.
inline uint32_t pcg_hash2(uint32_t key, uint32_t ctr) {
    uint32_t x = key * 747796405u + ctr * 2891336453u + 0x9e3779b9u;
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    uint32_t w = ((x >> ((x >> 28u) + 4u)) ^ x) * 277803737u;
    return (w >> 22u) ^ w;
}
// rand 0.9 StandardUniform for f32: 24 random bits * 2^-24
inline float u32_to_unit_float(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }

struct AuxRng {
    uint32_t key, ctr;
    float next() { return u32_to_unit_float(pcg_hash2(key, ctr++)); }
};
inline AuxRng aux_rng(uint32_t aux_base, uint32_t depth, uint32_t site) { return AuxRng{aux_base ^ ((depth * 4u + site) * 0x27d4eb2fu), 0u}; }

struct RandomSampler : SamplerBase {
    uint32_t seed, key = 0, ctr = 0, aux_key = 0;
    explicit RandomSampler(uint32_t seed_) : seed(seed_) {}
    void start_pixel_sample(uint32_t px, uint32_t py, uint32_t sample_index) override {
        key = (px * 0x9e3779b9u) ^ (py * 0x85ebca6bu) ^ (sample_index * 0xc2b2ae35u) ^ seed;
        ctr = 0;
        aux_key = key ^ 0x5bd1e995u;
    }
    float get_1d() override { return u32_to_unit_float(pcg_hash2(key, ctr++)); }
    Vec2 get_2d() override {
        Vec2 r;
        r.x = get_1d();
        r.y = get_1d();
        return r;
    }
    uint32_t aux_base() const override { return aux_key; }
};

}  // namespace orc
