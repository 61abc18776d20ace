// noise_kernels.cu — fBm fractal noise evaluator (hot loop 1).
//
// Replaces FractalJob<FractalGenerator<*Getter>, WriteTileData> (Noise/Fractal/Fractal.cs:20-138,
// getters :141-278, delegate table Noise/NoiseStage.cs:26-35).  The basis functions restate
// Unity.Mathematics.noise (com.unity.mathematics@1.2.1: snoise/cnoise/psrnoise/cellular), i.e. the
// webgl-noise algorithms, in the canonical evaluation order of oracle/noize_oracle.cpp.
//
// Shape of the kernel: FP32-pipe bound (about 2.3k ops per 4-byte store), no memory traffic but
// the store.  One thread owns NZ_CELLS consecutive-in-block cells (stride blockDim.x along x) so
// the independent octave chains of several cells interleave (ILP) and every store is a coalesced
// 128 B line per warp.  All hashing state lives in registers; there is no shared memory.
#include <math.h>
#include "nz_common.cuh"

namespace nz {
namespace {

// ---- Unity.Mathematics.noise common.cs ------------------------------------------------------
__device__ __forceinline__ float mod289(float x) { return fmaf(-floorf(x * (1.0f / 289.0f)), 289.0f, x); }
__device__ __forceinline__ float mod7(float x) { return fmaf(-floorf(x * (1.0f / 7.0f)), 7.0f, x); }
__device__ __forceinline__ float permute(float x) { return mod289(fmaf(34.0f, x, 1.0f) * x); }
__device__ __forceinline__ float taylorInvSqrt(float r) { return fmaf(-0.85373472095314f, r, 1.79284291400159f); }
__device__ __forceinline__ float fade(float t) { return t * t * t * fmaf(t, fmaf(t, 6.0f, -15.0f), 10.0f); }
__device__ __forceinline__ float fracf_(float x) { return x - floorf(x); }
__device__ __forceinline__ float lerpf_(float a, float b, float t) { return fmaf(t, b - a, a); }
__device__ __forceinline__ float stepf_(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }
__device__ __forceinline__ float dot2(float ax, float ay, float bx, float by) { return fmaf(ay, by, ax * bx); }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaf(az, bz, fmaf(ay, by, ax * bx));
}
__device__ __forceinline__ float dot4(float ax, float ay, float az, float aw, float bx, float by, float bz, float bw) {
    return fmaf(aw, bw, fmaf(az, bz, fmaf(ay, by, ax * bx)));
}
__device__ __forceinline__ float rectify(float v) { return (1.0f + v) * 0.5f; }

// ---- snoise(float2) ---------------------------------------------------------------------------
// Same operations as the oracle's snoise2; the only restructuring is exact: the middle corner's
// inner hash permute(iy + i1y) is one of the two hashes already computed for the outer corners
// (i1y is 0 or 1), so it is selected instead of recomputed.
//
// FRND (floorf) issues at 0.5 warp-inst/clk/SM on B200, 1/8 of the FP32 rate (profiles/r1_ubench_*), and
// the straightforward form needs 15 of them per octave, which saturates the XU pipe before the FMA pipe.
// Two more exact rewrites move 5 of them onto the FP32 pipe at no extra instruction count:
//   * the two INNER hashes use a round-to-nearest quotient (magic-number add) instead of floor: the
//     result r = u - 289*rn(u/289) is congruent to u mod 289 (|r| <= 144) and only ever feeds the outer
//     hash, whose polynomial (34x+1)x maps congruent inputs to congruent outputs, so the canonical
//     residue the outer hash produces is unchanged.  u <= 2.9e6 < 2^22 keeps the magic add exact.
//   * floor(gx + 0.5) == rn(gx) for every gx the 289 hash values can produce (no ties), checked
//     exhaustively in tests/test_oracle.py.
constexpr float NZ_MAGIC = 12582912.0f;  // 1.5 * 2^23: (x + MAGIC) - MAGIC == rint(x) for |x| < 2^22
__device__ __forceinline__ float permute_centered(float x) {
    const float u = fmaf(34.0f, x, 1.0f) * x;
    const float q = fmaf(u, 1.0f / 289.0f, NZ_MAGIC) - NZ_MAGIC;
    return fmaf(-289.0f, q, u);
}

// ---- per-CTA hash tables ------------------------------------------------------------------------------
// What a basis function derives from a final hash value p in [0,288] is a pure function of p:
//   simplex / perlin : gradient fold  x = 2*frac(p/41) - 1, h = |x| - 0.5, a0 = x - floor(x + 0.5)
//   cellular         : feature-point jitter  ox = frac(p/7) - 3/7, oy = mod7(floor(p/7))/7 - 3/7
//   psrnoise         : rotated gradient  (cos u, sin u), u = 2*pi*frac(p/41 + rot)
// so the kernel evaluates it once per CTA for all 289 values into shared memory with exactly the operations the
// straight-line code would use (same bits), and the octave loop replaces 8-10 instructions per corner (2-3 of them
// XU floors, or a sin and a cos) by a conversion and one LDS.64.
//
// Indexing.  FAST (the host proved every lattice index of the launch is below 2^21, so all hash arithmetic is
// exact integer arithmetic in float): every hash, the last one included, is kept as a CENTRED residue in
// [-144,144] (round-to-nearest quotient, no XU floor) and the table is indexed by residue + 144, which names the
// same canonical value.  Otherwise every hash keeps the canonical float-floor form, and the table is a CACHE of
// the pure function: a final hash value that is not an exact integer in [0,288] (mod289 of a coordinate beyond
// float's integer range returns 289, negatives or huge garbage) takes the straight-line code instead (a branch
// that is never taken for sane coordinates), so the result is bit-for-bit the oracle's everywhere.
constexpr int NZ_TAB = 289;
template <bool FAST>
__device__ __forceinline__ float hash_last(float x) { return FAST ? permute_centered(x) : permute(x); }
template <bool FAST>
__device__ __forceinline__ float hash_inner(float x) { return FAST ? permute_centered(x) : permute(x); }
template <bool FAST>
__device__ __forceinline__ float lattice_residue(float x) {
    if (FAST) return fmaf(-289.0f, fmaf(x, 1.0f / 289.0f, NZ_MAGIC) - NZ_MAGIC, x);
    return mod289(x);
}

__device__ __forceinline__ float2 gradient_fold(float p) {
    const float gx = fmaf(2.0f, fracf_(p * 0.024390243902439f), -1.0f);
    const float h = fabsf(gx) - 0.5f;
    const float a0 = gx - ((gx + NZ_MAGIC) - NZ_MAGIC);
    return make_float2(a0, h);
}
__device__ __forceinline__ float2 cellular_jitter(float p) {
    const float K = 0.142857142857f, Ko = 0.428571428571f;
    return make_float2(fracf_(p * K) - Ko, fmaf(mod7(floorf(p * K)), K, -Ko));
}
__device__ __forceinline__ float2 rotated_gradient(float p, float rot) {
    float u = fmaf(p, 0.0243902439f, rot);
    u = fracf_(u) * 6.28318530718f;
    return make_float2(cosf(u), sinf(u));
}
template <int TYPE>
__device__ __forceinline__ float2 table_function(float p) {
    if (TYPE == NZ_NOISE_CELLULAR) return cellular_jitter(p);
    if (TYPE == NZ_NOISE_PERIODIC_PERLIN) return rotated_gradient(p, 0.0f);
    if (TYPE == NZ_NOISE_ROTATED_SIMPLEX) return rotated_gradient(p, 0.62f);
    return gradient_fold(p);
}
template <int TYPE>
__device__ __forceinline__ void build_table(float2* tab) {
    for (int i = threadIdx.x; i < NZ_TAB; i += blockDim.x) {
        const int c = i - 144;
        const float p = (float)(c < 0 ? c + 289 : c);
        tab[i] = table_function<TYPE>(p);
    }
    __syncthreads();
}
template <int TYPE, bool FAST>
__device__ __forceinline__ float2 table_lookup(const float2* tab, float p) {
    const int i = __float2int_rn(p);
    if (FAST) return tab[i + 144];
    if ((unsigned)i <= 288u && (float)i == p) return tab[i > 144 ? i - 145 : i + 144];
    return table_function<TYPE>(p);
}
template <int TYPE>
__host__ __device__ constexpr bool uses_table() {
    return TYPE == NZ_NOISE_SIMPLEX || TYPE == NZ_NOISE_PERLIN || TYPE == NZ_NOISE_CELLULAR || TYPE == NZ_NOISE_PERIODIC_PERLIN ||
           TYPE == NZ_NOISE_ROTATED_SIMPLEX;
}

// returns dot(m, g), i.e. snoise(float2) / 130 (the getter folds the factor, see basis_value)
template <bool FAST>
__device__ __forceinline__ float snoise2_raw(float vx, float vy, const float2* gtab) {
    const float Cx = 0.211324865405187f, Cy = 0.366025403784439f;
    const float Cz = -0.577350269189626f;
    float s = dot2(vx, vy, Cy, Cy);
    float ix = floorf(vx + s), iy = floorf(vy + s);
    float t = dot2(ix, iy, Cx, Cx);
    float x0x = vx - ix + t, x0y = vy - iy + t;
    const bool xgty = x0x > x0y;
    float i1x = xgty ? 1.0f : 0.0f, i1y = xgty ? 0.0f : 1.0f;
    float x1x = x0x + Cx - i1x, x1y = x0y + Cx - i1y;
    float x2x = x0x + Cz, x2y = x0y + Cz;
    ix = lattice_residue<FAST>(ix);
    iy = lattice_residue<FAST>(iy);
    float py0 = permute_centered(iy), py1 = permute_centered(iy + 1.0f);
    const float hA = py0 + ix, hB = py1 + ix;     // (py + ix) + i1x with i1x in {0,1}: pick, do not recompute
    float p0 = hash_last<FAST>(hA);
    float p1 = hash_last<FAST>(xgty ? hA + 1.0f : hB);
    float p2 = hash_last<FAST>(hB + 1.0f);
    float m0 = fmaxf(fmaf(-x0y, x0y, fmaf(-x0x, x0x, 0.5f)), 0.0f);
    float m1 = fmaxf(fmaf(-x1y, x1y, fmaf(-x1x, x1x, 0.5f)), 0.0f);
    float m2 = fmaxf(fmaf(-x2y, x2y, fmaf(-x2x, x2x, 0.5f)), 0.0f);
    m0 = m0 * m0; m0 = m0 * m0;
    m1 = m1 * m1; m1 = m1 * m1;
    m2 = m2 * m2; m2 = m2 * m2;
    const float2 gr0 = table_lookup<NZ_NOISE_SIMPLEX, FAST>(gtab, p0), gr1 = table_lookup<NZ_NOISE_SIMPLEX, FAST>(gtab, p1),
                 gr2 = table_lookup<NZ_NOISE_SIMPLEX, FAST>(gtab, p2);
    const float a0 = gr0.x, h0 = gr0.y, a1 = gr1.x, h1 = gr1.y, a2 = gr2.x, h2 = gr2.y;
    m0 = m0 * taylorInvSqrt(fmaf(h0, h0, a0 * a0));
    m1 = m1 * taylorInvSqrt(fmaf(h1, h1, a1 * a1));
    m2 = m2 * taylorInvSqrt(fmaf(h2, h2, a2 * a2));
    float g0 = fmaf(h0, x0y, a0 * x0x);
    float g1 = fmaf(h1, x1y, a1 * x1x);
    float g2 = fmaf(h2, x2y, a2 * x2x);
    return dot3(m0, m1, m2, g0, g1, g2);
}

// ---- cnoise(float2) ---------------------------------------------------------------------------
template <bool FAST>
__device__ __forceinline__ float cnoise2(float Px, float Py, const float2* gtab) {
    float flx = floorf(Px), fly = floorf(Py);
    float pfx0 = Px - flx, pfy0 = Py - fly;
    float pfx1 = pfx0 - 1.0f, pfy1 = pfy0 - 1.0f;
    float pix0 = lattice_residue<FAST>(flx), pix1 = lattice_residue<FAST>(flx + 1.0f);
    float piy0 = lattice_residue<FAST>(fly), piy1 = lattice_residue<FAST>(fly + 1.0f);
    float hx0 = hash_inner<FAST>(pix0), hx1 = hash_inner<FAST>(pix1);   // inner hashes only feed the last hash
    float n[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float fx = (k & 1) ? pfx1 : pfx0, fy = (k >> 1) ? pfy1 : pfy0;
        float i = hash_last<FAST>(((k & 1) ? hx1 : hx0) + ((k >> 1) ? piy1 : piy0));
        const float2 g = table_lookup<NZ_NOISE_PERLIN, FAST>(gtab, i);      // (gx, gy) = (g - floor(g + 0.5), |g| - 0.5), g = 2*frac(i/41) - 1
        float norm = taylorInvSqrt(dot2(g.x, g.y, g.x, g.y));
        n[k] = dot2(g.x * norm, g.y * norm, fx, fy);
    }
    float fdx = fade(pfx0), fdy = fade(pfy0);
    return 2.3f * lerpf_(lerpf_(n[0], n[1], fdx), lerpf_(n[2], n[3], fdx), fdy);
}

// ---- psrnoise(float2, float2 per, float rot) --------------------------------------------------
// The hash arguments are half-integers that can exceed float's exact-product range, so the hashes keep the
// canonical float-floor form; only the (cos, sin) of the final hash value comes from the table.
// FAST (wrapped coordinates are exact half-integers, |px| <= 1061, |py| < 102): both hashes use the round-to-nearest
// residue.  The inner product (34x+1)x exceeds 2^24 and is rounded, but both residue forms subtract an exact multiple of
// 289 from the SAME rounded product, so they are congruent; the outer hash argument is then a small exact integer and
// its centred residue indexes the table directly (no floor, no range check).
template <int TYPE, bool FAST>
__device__ __forceinline__ void rgrad2(float px, float py, const float2* rtab, float& gx, float& gy) {
    float2 g;
    if (FAST) g = rtab[__float2int_rn(permute_centered(permute_centered(px) + py)) + 144];
    else g = table_lookup<TYPE, false>(rtab, permute(permute(px) + py));
    gx = g.x;
    gy = g.y;
}
// fmodf(p, per) for |p| < 2^22 (FAST: the host proved it): fmod is exact, so p - per*trunc(p/per) with a one-step
// correction of the estimated quotient gives the same value as the library routine (whose generic bit-serial loop
// costs ~80 instructions at C5 coordinates).  The sign of a zero result may differ from fmodf's; it cannot reach the
// output: permute() maps +-0 to +0 and every other use adds the value to a non-zero or to another hash input.
template <bool FAST>
__device__ __forceinline__ float fmod_period(float p, float per, float inv_per) {
    if (!FAST) return fmodf(p, per);
    const float a = fabsf(p);
    const float q = truncf(a * inv_per);
    float r = fmaf(-per, q, a);             // exact: |r| < 2*per, a and per*q below 2^23
    r = r < 0.0f ? r + per : r;
    r = r >= per ? r - per : r;
    return copysignf(r, p);
}

template <int TYPE, bool FAST>
__device__ __forceinline__ float psrnoise2(float posx, float posy, float perx, float pery, const float2* rtab) {
    posy += 0.001f;
    float ux = fmaf(posy, 0.5f, posx), uy = posy;
    float i0x = floorf(ux), i0y = floorf(uy);
    float f0x = ux - i0x, f0y = uy - i0y;
    const bool c = f0x > f0y;
    float i1x = c ? 1.0f : 0.0f, i1y = c ? 0.0f : 1.0f;
    float p0x = fmaf(-i0y, 0.5f, i0x), p0y = i0y;
    float p1x = p0x + i1x - i1y * 0.5f, p1y = p0y + i1y;
    float p2x = p0x + 0.5f, p2y = p0y + 1.0f;
    float d0x = posx - p0x, d0y = posy - p0y;
    float d1x = posx - p1x, d1y = posy - p1y;
    float d2x = posx - p2x, d2y = posy - p2y;
    const float ipx = 1.0f / perx, ipy = 1.0f / pery;      // compile-time constants at both call sites
    float xw0 = fmod_period<FAST>(p0x, perx, ipx), yw0 = fmod_period<FAST>(p0y, pery, ipy);
    float xw1, xw2, yw1, yw2;
    if (FAST && p0x >= 1.0f && p0y >= 0.0f) {
        // All three corners are non-negative here (p1x >= p0x - 0.5, p1y >= p0y, p2 = p0 + (0.5, 1)), every value is a
        // multiple of 0.5 below 2^22 (exact), and the corners differ by less than one period, so the other two residues
        // are the first one plus the offset with ONE wrap: the same values fmod gives, for 11 instructions instead of 40.
        yw2 = yw0 + 1.0f;
        yw2 = yw2 >= pery ? yw2 - pery : yw2;
        yw1 = c ? yw0 : yw2;                                  // i1y = c ? 0 : 1
        xw2 = xw0 + 0.5f;
        xw2 = xw2 >= perx ? xw2 - perx : xw2;
        xw1 = xw0 + (c ? 1.0f : -0.5f);                       // i1x - 0.5*i1y
        xw1 = xw1 >= perx ? xw1 - perx : xw1;
        xw1 = xw1 < 0.0f ? xw1 + perx : xw1;
    } else {
        xw1 = fmod_period<FAST>(p1x, perx, ipx); xw2 = fmod_period<FAST>(p2x, perx, ipx);
        yw1 = fmod_period<FAST>(p1y, pery, ipy); yw2 = fmod_period<FAST>(p2y, pery, ipy);
    }
    float g0x, g0y, g1x, g1y, g2x, g2y;
    rgrad2<TYPE, FAST>(fmaf(0.5f, yw0, xw0), yw0, rtab, g0x, g0y);
    rgrad2<TYPE, FAST>(fmaf(0.5f, yw1, xw1), yw1, rtab, g1x, g1y);
    rgrad2<TYPE, FAST>(fmaf(0.5f, yw2, xw2), yw2, rtab, g2x, g2y);
    float w0 = dot2(g0x, g0y, d0x, d0y), w1 = dot2(g1x, g1y, d1x, d1y), w2 = dot2(g2x, g2y, d2x, d2y);
    float t0 = fmaxf(0.8f - dot2(d0x, d0y, d0x, d0y), 0.0f);
    float t1 = fmaxf(0.8f - dot2(d1x, d1y, d1x, d1y), 0.0f);
    float t2 = fmaxf(0.8f - dot2(d2x, d2y, d2x, d2y), 0.0f);
    t0 = t0 * t0; t0 = t0 * t0;
    t1 = t1 * t1; t1 = t1 * t1;
    t2 = t2 * t2; t2 = t2 * t2;
    return 11.0f * dot3(t0, t1, t2, w0, w1, w2);
}

// ---- cellular(float2) -> F1*F2 rectified ------------------------------------------------------
template <bool FAST>
__device__ __forceinline__ float cellular2_rectified(float Px, float Py, const float2* jtab) {
    float flx = floorf(Px), fly = floorf(Py);
    float Pix = lattice_residue<FAST>(flx), Piy = lattice_residue<FAST>(fly);
    float Pfx = Px - flx, Pfy = Py - fly;
    float d[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float oic = (float)(c - 1);
        const float xo = 0.5f - (float)c;
        float pxc = hash_inner<FAST>(Pix + oic);   // inner hash: only feeds the last hash
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float oij = (float)(j - 1), ofj = (float)j - 0.5f;
            const float2 o = table_lookup<NZ_NOISE_CELLULAR, FAST>(jtab, hash_last<FAST>(pxc + Piy + oij));   // (ox, oy)
            float dx = Pfx + xo + o.x;
            float dy = Pfy - ofj + o.y;
            d[c][j] = fmaf(dy, dy, dx * dx);
        }
    }
    float d1[3], d2[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        float d1a = fminf(d[0][j], d[1][j]);
        float t = fmaxf(d[0][j], d[1][j]);
        t = fminf(t, d[2][j]);
        d1[j] = fminf(d1a, t);
        d2[j] = fmaxf(d1a, t);
    }
    if (!(d1[0] < d1[1])) { float t = d1[0]; d1[0] = d1[1]; d1[1] = t; }
    if (!(d1[0] < d1[2])) { float t = d1[0]; d1[0] = d1[2]; d1[2] = t; }
    d1[1] = fminf(d1[1], d2[1]);
    d1[2] = fminf(d1[2], d2[2]);
    d1[1] = fminf(d1[1], d1[2]);
    d1[1] = fminf(d1[1], d2[0]);
    return rectify(sqrtf(d1[0])) * rectify(sqrtf(d1[1]));
}

// ---- snoise(float3) ---------------------------------------------------------------------------
// The normalised corner gradient is a pure function of the final hash value p in [0,288]; simplex3_gradient() is the
// straight-line code (it costs 5 floors per corner).  FAST (host-checked lattice range): the kernel evaluates it once
// per CTA into a 289-entry float4 table and every hash of the chain keeps a centred residue (no floor), as in 2-D.
__device__ __forceinline__ float4 simplex3_gradient(float p) {
    const float n_ = 0.142857142857f;
    const float nsx = n_ * 2.0f - 0.0f, nsy = n_ * 0.5f - 1.0f, nsz = n_ * 1.0f - 0.0f;
    float j = fmaf(-49.0f, floorf(p * nsz * nsz), p);
    float x_ = floorf(j * nsz);
    float y_ = floorf(fmaf(-7.0f, x_, j));
    float X = fmaf(x_, nsx, nsy), Y = fmaf(y_, nsx, nsy);
    float H = 1.0f - fabsf(X) - fabsf(Y);
    float sh = -stepf_(H, 0.0f);
    float sx = fmaf(floorf(X), 2.0f, 1.0f), sy = fmaf(floorf(Y), 2.0f, 1.0f);
    float Px = fmaf(sx, sh, X), Py = fmaf(sy, sh, Y), Pz = H;
    float norm = taylorInvSqrt(dot3(Px, Py, Pz, Px, Py, Pz));
    return make_float4(Px * norm, Py * norm, Pz * norm, 0.0f);
}
template <bool FAST>
__device__ float snoise3(float vx, float vy, float vz, const float4* gtab3) {
    const float Cx = 1.0f / 6.0f, Cy = 1.0f / 3.0f;
    float s = dot3(vx, vy, vz, Cy, Cy, Cy);
    float i0 = floorf(vx + s), i1_ = floorf(vy + s), i2_ = floorf(vz + s);
    float t = dot3(i0, i1_, i2_, Cx, Cx, Cx);
    float x0[3] = {vx - i0 + t, vy - i1_ + t, vz - i2_ + t};
    float g[3] = {stepf_(x0[1], x0[0]), stepf_(x0[2], x0[1]), stepf_(x0[0], x0[2])};
    float l[3] = {1.0f - g[0], 1.0f - g[1], 1.0f - g[2]};
    float i1[3] = {fminf(g[0], l[2]), fminf(g[1], l[0]), fminf(g[2], l[1])};
    float i2[3] = {fmaxf(g[0], l[2]), fmaxf(g[1], l[0]), fmaxf(g[2], l[1])};
    float xs[4][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        xs[0][k] = x0[k];
        xs[1][k] = x0[k] - i1[k] + Cx;
        xs[2][k] = x0[k] - i2[k] + Cy;
        xs[3][k] = x0[k] - 0.5f;
    }
    float ii[3] = {lattice_residue<FAST>(i0), lattice_residue<FAST>(i1_), lattice_residue<FAST>(i2_)};
    const float oz[4] = {0.0f, i1[2], i2[2], 1.0f}, oy[4] = {0.0f, i1[1], i2[1], 1.0f}, ox[4] = {0.0f, i1[0], i2[0], 1.0f};
    float m[4], pd[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float p = hash_last<FAST>(hash_inner<FAST>(hash_inner<FAST>(ii[2] + oz[k]) + ii[1] + oy[k]) + ii[0] + ox[k]);
        const float4 G = FAST ? gtab3[__float2int_rn(p) + 144] : simplex3_gradient(p);
        float mm = fmaxf(0.6f - dot3(xs[k][0], xs[k][1], xs[k][2], xs[k][0], xs[k][1], xs[k][2]), 0.0f);
        mm = mm * mm;
        m[k] = mm * mm;
        pd[k] = dot3(G.x, G.y, G.z, xs[k][0], xs[k][1], xs[k][2]);
    }
    return 42.0f * dot4(m[0], m[1], m[2], m[3], pd[0], pd[1], pd[2], pd[3]);
}

// ---- cnoise(float3) ---------------------------------------------------------------------------
// Same scheme: the normalised gradient of a corner is a pure function of its final hash value.
__device__ __forceinline__ float4 perlin3_gradient(float ixyz) {
    float gx = ixyz * (1.0f / 7.0f);
    float gy = fracf_(floorf(gx) * (1.0f / 7.0f)) - 0.5f;
    gx = fracf_(gx);
    float gz = 0.5f - fabsf(gx) - fabsf(gy);
    float sz = stepf_(gz, 0.0f);
    gx = gx - sz * (stepf_(0.0f, gx) - 0.5f);
    gy = gy - sz * (stepf_(0.0f, gy) - 0.5f);
    float norm = taylorInvSqrt(dot3(gx, gy, gz, gx, gy, gz));
    return make_float4(gx * norm, gy * norm, gz * norm, 0.0f);
}
template <bool FAST>
__device__ float cnoise3(float Px, float Py, float Pz, const float4* gtab3) {
    float P[3] = {Px, Py, Pz};
    float Pi0[3], Pi1[3], Pf0[3], Pf1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float fl = floorf(P[k]);
        Pi0[k] = lattice_residue<FAST>(fl);
        Pi1[k] = lattice_residue<FAST>(fl + 1.0f);
        Pf0[k] = P[k] - fl;
        Pf1[k] = Pf0[k] - 1.0f;
    }
    float n[2][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float ix = (k & 1) ? Pi1[0] : Pi0[0], iy = (k >> 1) ? Pi1[1] : Pi0[1];
        float ixy = hash_inner<FAST>(hash_inner<FAST>(ix) + iy);
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
            float ixyz = hash_last<FAST>(ixy + (sl ? Pi1[2] : Pi0[2]));
            const float4 G = FAST ? gtab3[__float2int_rn(ixyz) + 144] : perlin3_gradient(ixyz);
            float fx = (k & 1) ? Pf1[0] : Pf0[0], fy = (k >> 1) ? Pf1[1] : Pf0[1], fz = sl ? Pf1[2] : Pf0[2];
            n[sl][k] = dot3(G.x, G.y, G.z, fx, fy, fz);
        }
    }
    float fdx = fade(Pf0[0]), fdy = fade(Pf0[1]), fdz = fade(Pf0[2]);
    float nz0 = lerpf_(n[0][0], n[1][0], fdz), nz1 = lerpf_(n[0][1], n[1][1], fdz);
    float nz2 = lerpf_(n[0][2], n[1][2], fdz), nz3 = lerpf_(n[0][3], n[1][3], fdz);
    return 2.2f * lerpf_(lerpf_(nz0, nz2, fdy), lerpf_(nz1, nz3, fdy), fdx);
}

// ---- basis getters, Fractal.cs:141-278 ----------------------------------------------------------
template <int TYPE, bool FAST>
__device__ __forceinline__ float basis_value(float x, float z, const float2* gtab, const float4* gtab3) {
    if (TYPE == NZ_NOISE_SIN) {
        float vx = fmaf(0.5f, sinf(x), 0.5f), vz = fmaf(0.5f, sinf(z), 0.5f);
        return vx * vz;
    } else if (TYPE == NZ_NOISE_PERLIN) {
        return rectify(cnoise2<FAST>(x, z, gtab));
    } else if (TYPE == NZ_NOISE_PERIODIC_PERLIN) {
        return rectify(psrnoise2<TYPE, FAST>(x, z, 1010.0f, 102.0f, gtab));
    } else if (TYPE == NZ_NOISE_SIMPLEX) {
        return fmaf(65.0f, snoise2_raw<FAST>(x, z, gtab), 0.5f);  // Rectify(130*d) = (1 + 130*d)/2 = 0.5 + 65*d
    } else if (TYPE == NZ_NOISE_ROTATED_SIMPLEX) {
        return rectify(psrnoise2<TYPE, FAST>(x, z, 1010.0f, 102.0f, gtab));
    } else if (TYPE == NZ_NOISE_CELLULAR) {
        return cellular2_rectified<FAST>(x, z, gtab);
    } else {
        float xz = x + z;
        float s2 = xz * -0.211324865405187f;
        float xr = x + s2, zr = z + s2;
        float yr = xz * -0.577350269189626f;
        return rectify(TYPE == NZ_NOISE_DOMAIN_ROTATED_PERLIN ? cnoise3<FAST>(xr, zr, yr, gtab3) : snoise3<FAST>(xr, zr, yr, gtab3));
    }
}

// ---- the fBm kernel: FractalGenerator.NoiseValue / Execute, Fractal.cs:114-138 -------------------
constexpr int NZ_FBM_THREADS = 128;

template <int TYPE, int CELLS, bool FAST>
__global__ void __launch_bounds__(NZ_FBM_THREADS) fbm_kernel(float* __restrict__ dst, FractalParams p) {
    __shared__ float2 gtab[uses_table<TYPE>() ? NZ_TAB : 1];
    if (uses_table<TYPE>()) build_table<TYPE>(gtab);
    constexpr bool TAB3 = FAST && (TYPE == NZ_NOISE_DOMAIN_ROTATED_PERLIN || TYPE == NZ_NOISE_DOMAIN_ROTATED_SIMPLEX);
    __shared__ float4 gtab3[TAB3 ? NZ_TAB : 1];
    if (TAB3) {
        for (int i = threadIdx.x; i < NZ_TAB; i += blockDim.x) {
            const int c = i - 144;
            const float pv = (float)(c < 0 ? c + 289 : c);
            gtab3[i] = TYPE == NZ_NOISE_DOMAIN_ROTATED_PERLIN ? perlin3_gradient(pv) : simplex3_gradient(pv);
        }
        __syncthreads();
    }
    const int r = blockIdx.y;
    const int xbase = blockIdx.x * (NZ_FBM_THREADS * CELLS) + threadIdx.x;
    const float zi = ((float)(p.z_first + r) + p.posz) / p.noise_size;
    float xi[CELLS], t[CELLS];
#pragma unroll
    for (int c = 0; c < CELLS; c++) {
        xi[c] = ((float)(xbase + c * NZ_FBM_THREADS) + p.posx) / p.noise_size;
        t[c] = 0.0f;
    }
    float detune = 0.0f, f = 1.0f, a = p.start_amp;
    for (int i = 0; i < p.octaves; i++) {
        const float zV = f * zi;
#pragma unroll
        for (int c = 0; c < CELLS; c++) t[c] = fmaf(a, basis_value<TYPE, FAST>(f * xi[c], zV, gtab, gtab3), t[c]);
        detune += p.detune_rate;
        f *= (p.stepdown - detune);
        a *= p.G;
    }
    float* row = dst + (size_t)r * p.width;
#pragma unroll
    for (int c = 0; c < CELLS; c++) {
        const int x = xbase + c * NZ_FBM_THREADS;
        if (x < p.width) row[x] = t[c] / p.norm;
    }
}

template <int TYPE, int CELLS>
int32_t launch_typed(float* d_dst, const FractalParams& p, cudaStream_t s) {
    dim3 grid(cdiv(p.width, NZ_FBM_THREADS * CELLS), p.rows);
    const bool is3d = TYPE == NZ_NOISE_DOMAIN_ROTATED_PERLIN || TYPE == NZ_NOISE_DOMAIN_ROTATED_SIMPLEX;
    if ((TYPE == NZ_NOISE_SIMPLEX || TYPE == NZ_NOISE_PERLIN || TYPE == NZ_NOISE_CELLULAR || TYPE == NZ_NOISE_PERIODIC_PERLIN ||
         TYPE == NZ_NOISE_ROTATED_SIMPLEX || is3d) && (is3d ? p.fast_hash3d : p.fast_hash))
        fbm_kernel<TYPE, CELLS, true><<<grid, NZ_FBM_THREADS, 0, s>>>(d_dst, p);
    else
        fbm_kernel<TYPE, CELLS, false><<<grid, NZ_FBM_THREADS, 0, s>>>(d_dst, p);
    NZ_LAUNCHED();
    return NZ_OK;
}

}  // namespace

int32_t launch_fractal(float* d_dst, int noise_type, const FractalParams& p, cudaStream_t s) {
    // simplex / cellular on a window large enough to fill the GPU: packed-pair kernels (fbmpair_kernels.cu), bit-identical
    // to the scalar kernel below; NZ_FBM_PATH=scalar|pair forces one (tests compare the two bitwise)
    {
        const char* force = getenv("NZ_FBM_PATH");
        const bool scalar = force && force[0] == 's', pair = force && force[0] == 'p';
        if (!scalar && fractal_pair_possible(noise_type, p) && (pair || fractal_pair_supported(noise_type, p)))
            return launch_fractal_pair(d_dst, noise_type, p, s);
    }
    if (p.rows > 65535) {
        // gridDim.y limit: split into row chunks
        FractalParams q = p;
        for (int r0 = 0; r0 < p.rows; r0 += 32768) {
            q.rows = p.rows - r0 < 32768 ? p.rows - r0 : 32768;
            q.z_first = p.z_first + r0;
            int32_t rc = launch_fractal(d_dst + (size_t)r0 * p.width, noise_type, q, s);
            if (rc != NZ_OK) return rc;
        }
        return NZ_OK;
    }
    switch (noise_type) {
        case NZ_NOISE_SIN: return launch_typed<NZ_NOISE_SIN, 2>(d_dst, p, s);
        case NZ_NOISE_PERLIN: return launch_typed<NZ_NOISE_PERLIN, 2>(d_dst, p, s);
        case NZ_NOISE_PERIODIC_PERLIN: return launch_typed<NZ_NOISE_PERIODIC_PERLIN, 1>(d_dst, p, s);
        case NZ_NOISE_SIMPLEX: {
            // cells per thread is a tuning knob (NZ_FBM_CELLS=1|2|4 overrides the default while profiling)
            static const int cells = [] {
                const char* e = getenv("NZ_FBM_CELLS");
                return e ? atoi(e) : 4;
            }();
            if (cells == 1) return launch_typed<NZ_NOISE_SIMPLEX, 1>(d_dst, p, s);
            if (cells == 4) return launch_typed<NZ_NOISE_SIMPLEX, 4>(d_dst, p, s);
            return launch_typed<NZ_NOISE_SIMPLEX, 2>(d_dst, p, s);
        }
        case NZ_NOISE_ROTATED_SIMPLEX: return launch_typed<NZ_NOISE_ROTATED_SIMPLEX, 1>(d_dst, p, s);
        case NZ_NOISE_CELLULAR: return launch_typed<NZ_NOISE_CELLULAR, 1>(d_dst, p, s);
        case NZ_NOISE_DOMAIN_ROTATED_PERLIN: return launch_typed<NZ_NOISE_DOMAIN_ROTATED_PERLIN, 1>(d_dst, p, s);
        case NZ_NOISE_DOMAIN_ROTATED_SIMPLEX: return launch_typed<NZ_NOISE_DOMAIN_ROTATED_SIMPLEX, 1>(d_dst, p, s);
    }
    set_error("nz_fractal: noise_type %d out of range", noise_type);
    return NZ_E_INVALID;
}

}  // namespace nz
