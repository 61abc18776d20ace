// fbmpair_kernels.cu — simplex fBm, second generation: packed FP32 pairs + bank-private hash tables.
//
// Same function as fbm_kernel<NZ_NOISE_SIMPLEX, *, FAST> in noise_kernels.cu (FractalGenerator<SimplexGetter>,
// Noise/Fractal/Fractal.cs:114-138,196-205) and bit-for-bit the same results (tests compare the two kernels
// bitwise); what changes is how the B200 executes it.  The first-generation kernel is issue-bound: 111
// instructions per cell and octave, 90 of them scalar FP32, 92 % issue-slot utilisation.  Two changes:
//
//  1. Packed FP32 (FFMA2/FMUL2, sm_100).  A thread owns 4 cells of one column (4 consecutive rows) and carries
//     them as two float2 pairs; every FP operation of the octave body is one f32x2 instruction per pair.  An
//     f32x2 instruction occupies the FP32 pipe for two cycles but only ONE issue slot (tools/ubench2: FFMA2
//     1.92 warp-inst/clk/SM, and ALU/LSU instructions fill the freed slots), so the FP work costs half the issue
//     bandwidth.  Each half is an IEEE round-to-nearest fma/mul/add, identical to the scalar instruction.
//
//  2. The hash becomes three table walks.  With every lattice index below 2^21 (FAST, host-checked) the hash
//     chain  p = permute(permute(iy + j) + ix + i)  is exact integer arithmetic, so
//        T1[iy mod 289]            = permute(iy)                       (one LDS per lattice row)
//        T2[(T1 + ix) mod 289]     = (a0, h, norm) of permute(T1 + ix) (three LDS per corner)
//     replaces 5 float hashes (25 FP instructions), the gradient fold and the normalisation polynomial
//     (9 more).  To make random lookups free of bank conflicts each table is BANK-PRIVATE: 32 copies, one per
//     lane, interleaved so that lane l only ever touches bank l (row stride 128 B).  Four tables of 290 rows
//     = 148,480 B of shared memory — one persistent 1024-thread CTA per SM, which a B200 SM (227 KB) holds.
//     Indices never leave the float domain through a conversion: the residue r is turned into address bits by
//     the magic add (r + 1.5*2^23 has r in its low mantissa bits) and one shift-add; the table entries carry
//     the compensating constants.  The wrap of (T1 + ix) into [0,289) is an unsigned min.
//
// Result: ~57 issue slots and ~52 FP-pipe cycles per cell and octave instead of 111 / 90.
#include <stdlib.h>
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr float MAGIC = 12582912.0f;                                   // 1.5 * 2^23
constexpr uint32_t MAGIC_SHL7 = (uint32_t)(0x4B400000ull << 7);        // bits(MAGIC) << 7, mod 2^32
constexpr int PAIR_THREADS = 1024;
constexpr int ROW_BYTES = 128;                                         // 32 lanes x 4 B
constexpr int TAB_ROWS = 290;                                          // 289 residues + one wrapped row
constexpr int TAB_BYTES = TAB_ROWS * ROW_BYTES;
constexpr int OFF_T1 = 0, OFF_A = TAB_BYTES, OFF_H = 2 * TAB_BYTES, OFF_N = 3 * TAB_BYTES;
constexpr int PAIR_SMEM = 4 * TAB_BYTES;
constexpr uint32_t WRAP = 289u * ROW_BYTES;

typedef float2 P;
__device__ __forceinline__ P bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ P neg(P a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ P pfma(P a, P b, P c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ P pmul(P a, P b) { return __fmul2_rn(a, b); }
// Additions are issued as fma(a, 1, b), which is the same IEEE sum: ptxas 12.9 contracts a mul.rn.f32x2 feeding
// an add.rn.f32x2 into one FFMA2 even under --fmad=false, which would change the rounding the oracle fixes.
__device__ __forceinline__ P padd(P a, P b) { return __ffma2_rn(a, bc(1.0f), b); }
__device__ __forceinline__ P psub(P a, P b) { return __ffma2_rn(a, bc(1.0f), neg(b)); }
__device__ __forceinline__ P pfloor(P a) { return make_float2(floorf(a.x), floorf(a.y)); }
__device__ __forceinline__ P pmax0(P a) { return make_float2(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)); }

// table reads: byte offsets relative to the start of the dynamic shared memory (the CTA's window base rides in a
// uniform register of the LDS, so offsets can be wrapped with an unsigned min)
extern __shared__ __align__(16) unsigned char sm[];
__device__ __forceinline__ uint32_t lds_u32(uint32_t off) { return *reinterpret_cast<const uint32_t*>(sm + off); }
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t off) { return *reinterpret_cast<const float*>(sm + OFF + off); }

// canonical hash of an integer residue class, exact (the float form is exact on this domain: tests/test_oracle.py)
__device__ __forceinline__ int permute_int(int a) {
    a %= 289;
    if (a < 0) a += 289;
    return ((34 * a + 1) * a) % 289;
}

// Every lane writes its own copy (its own bank) of every row.
__device__ void build_tables() {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int r = warp; r < TAB_ROWS; r += nwarps) {
        // T1 row r  <->  lattice residue iy = r - 144 in [-144, 145]:
        //   byte offset of T2 row (permute(iy) + 144), this lane's column, minus the bits the magic-number form
        //   of ix adds when it is shifted in (see the octave body)
        const uint32_t t1 = (uint32_t)(permute_int(r - 144) + 144) * ROW_BYTES + lane * 4 - MAGIC_SHL7;
        *reinterpret_cast<uint32_t*>(sm + OFF_T1 + r * ROW_BYTES + lane * 4) = t1;
        // T2 row r  <->  hash argument (r - 144) mod 289 (row 289 repeats row 0): the gradient fold of its hash,
        // with the operations of gradient_fold()/taylorInvSqrt in noise_kernels.cu
        const float p = (float)permute_int(r - 144);
        const float fr = p * 0.024390243902439f;
        const float gx = fmaf(2.0f, fr - floorf(fr), -1.0f);
        const float h = fabsf(gx) - 0.5f;
        const float a0 = gx - ((gx + MAGIC) - MAGIC);
        const float n = fmaf(-0.85373472095314f, fmaf(h, h, a0 * a0), 1.79284291400159f);
        *reinterpret_cast<float*>(sm + OFF_A + r * ROW_BYTES + lane * 4) = a0;
        *reinterpret_cast<float*>(sm + OFF_H + r * ROW_BYTES + lane * 4) = h;
        *reinterpret_cast<float*>(sm + OFF_N + r * ROW_BYTES + lane * 4) = n;
    }
    __syncthreads();
}

struct Corner {
    float a0, h, n;
};
__device__ __forceinline__ Corner fetch(uint32_t addr) {
    Corner c;
    c.a0 = lds_f32<OFF_A>(addr);
    c.h = lds_f32<OFF_H>(addr);
    c.n = lds_f32<OFF_N>(addr);
    return c;
}

// one octave of two cells (a pair) that share vx; returns dot(m, g) = snoise/130 for both
__device__ __forceinline__ P snoise_pair(float vx, float vxCy, P vy, uint32_t c1) {
    const float Cx = 0.211324865405187f, Cy = 0.366025403784439f, Cz = -0.577350269189626f;
    const P s = pfma(vy, bc(Cy), bc(vxCy));                       // dot2(vx, vy, Cy, Cy)
    const P ix = pfloor(padd(bc(vx), s)), iy = pfloor(padd(vy, s));
    const P t = pfma(iy, bc(Cx), pmul(ix, bc(Cx)));               // dot2(ix, iy, Cx, Cx)
    const P x0x = padd(psub(bc(vx), ix), t), x0y = padd(psub(vy, iy), t);
    const bool gA = x0x.x > x0y.x, gB = x0x.y > x0y.y;
    const P i1x = make_float2(gA ? 1.0f : 0.0f, gB ? 1.0f : 0.0f);
    const P i1y = make_float2(gA ? 0.0f : 1.0f, gB ? 0.0f : 1.0f);
    const P x1x = psub(padd(x0x, bc(Cx)), i1x), x1y = psub(padd(x0y, bc(Cx)), i1y);
    const P x2x = padd(x0x, bc(Cz)), x2y = padd(x0y, bc(Cz));
    // lattice residues (centred, exact) and their address bits
    const P rx = pfma(bc(-289.0f), psub(pfma(ix, bc(1.0f / 289.0f), bc(MAGIC)), bc(MAGIC)), ix);
    const P ry = pfma(bc(-289.0f), psub(pfma(iy, bc(1.0f / 289.0f), bc(MAGIC)), bc(MAGIC)), iy);
    const P bx = padd(rx, bc(MAGIC)), by = padd(ry, bc(MAGIC));
    Corner c0[2], c1_[2], c2[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const uint32_t ixb = __float_as_uint(e ? bx.y : bx.x) << 7, iyb = __float_as_uint(e ? by.y : by.x) << 7;
        const bool g = e ? gB : gA;
        const uint32_t a1 = iyb + c1;                              // T1 row of iy (this lane's column)
        const uint32_t pv0 = lds_u32(a1), pv1 = lds_u32(a1 + ROW_BYTES);
        uint32_t hA = ixb + pv0, hB = ixb + pv1;                   // T2 row (permute(iy+j) + ix + 144), unwrapped
        hA = min(hA, hA - WRAP);                                   // row index into [0,289): unsigned min
        hB = min(hB, hB - WRAP);
        const uint32_t hM = g ? hA + ROW_BYTES : hB;               // middle corner: (i1x, i1y) = (1,0) or (0,1)
        c0[e] = fetch(hA);
        c1_[e] = fetch(hM);
        c2[e] = fetch(hB + ROW_BYTES);
    }
    P m0 = pmax0(pfma(neg(x0y), x0y, pfma(neg(x0x), x0x, bc(0.5f))));
    P m1 = pmax0(pfma(neg(x1y), x1y, pfma(neg(x1x), x1x, bc(0.5f))));
    P m2 = pmax0(pfma(neg(x2y), x2y, pfma(neg(x2x), x2x, bc(0.5f))));
    m0 = pmul(m0, m0); m0 = pmul(m0, m0);
    m1 = pmul(m1, m1); m1 = pmul(m1, m1);
    m2 = pmul(m2, m2); m2 = pmul(m2, m2);
    m0 = pmul(m0, make_float2(c0[0].n, c0[1].n));
    m1 = pmul(m1, make_float2(c1_[0].n, c1_[1].n));
    m2 = pmul(m2, make_float2(c2[0].n, c2[1].n));
    const P g0 = pfma(make_float2(c0[0].h, c0[1].h), x0y, pmul(make_float2(c0[0].a0, c0[1].a0), x0x));
    const P g1 = pfma(make_float2(c1_[0].h, c1_[1].h), x1y, pmul(make_float2(c1_[0].a0, c1_[1].a0), x1x));
    const P g2 = pfma(make_float2(c2[0].h, c2[1].h), x2y, pmul(make_float2(c2[0].a0, c2[1].a0), x2x));
    return pfma(m2, g2, pfma(m1, g1, pmul(m0, g0)));              // dot3(m, g)
}

// Work item = 4*(PAIR_THREADS >> wshift) rows x (1 << wshift) columns; thread (ty, tx) owns rows 4*ty..4*ty+3 of
// column tx.  Lanes are consecutive columns, so every row store is a coalesced 128 B line per warp.
__global__ void __launch_bounds__(PAIR_THREADS, 1) fbm_simplex_pair_kernel(float* __restrict__ dst, FractalParams p, int wshift,
                                                                          int col_blocks, int n_items) {
    build_tables();
    const int lane = threadIdx.x & 31;
    uint32_t c1 = OFF_T1 + 144 * ROW_BYTES + lane * 4 - MAGIC_SHL7;
    asm volatile("" : "+r"(c1));   // keep the sum in one register (the constant does not fit an LDS immediate)
    const int tx = threadIdx.x & ((1 << wshift) - 1), ty = threadIdx.x >> wshift;
    const int rows_per_item = 4 * (PAIR_THREADS >> wshift);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cb = item % col_blocks, rg = item / col_blocks;
        const int x = (cb << wshift) + tx;
        const int r0 = rg * rows_per_item + 4 * ty;
        const float xi = ((float)x + p.posx) / p.noise_size;
        P zi[2], t[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            zi[q].x = ((float)(p.z_first + r0 + 2 * q) + p.posz) / p.noise_size;
            zi[q].y = ((float)(p.z_first + r0 + 2 * q + 1) + p.posz) / p.noise_size;
            t[q] = bc(0.0f);
        }
        float detune = 0.0f, f = 1.0f, a = p.start_amp;
        for (int i = 0; i < p.octaves; i++) {
            const float vx = f * xi;
            const float vxCy = vx * 0.366025403784439f;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const P raw = snoise_pair(vx, vxCy, pmul(bc(f), zi[q]), c1);
                t[q] = pfma(bc(a), pfma(bc(65.0f), raw, bc(0.5f)), t[q]);   // a * Rectify(130 * raw) + t
            }
            detune += p.detune_rate;
            f *= (p.stepdown - detune);
            a *= p.G;
        }
        if (x < p.width) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int r = r0 + 2 * q;
                if (r < p.rows) dst[(size_t)r * p.width + x] = t[q].x / p.norm;
                if (r + 1 < p.rows) dst[(size_t)(r + 1) * p.width + x] = t[q].y / p.norm;
            }
        }
    }
}

// =====================================================================================================================
// cellular (Worley F1*F2) fBm, same two techniques.  cellular2D hashes the 3x3 neighbourhood of the lattice cell:
//   p(c, j) = permute(permute(Pix + c-1) + Piy + j-1)  ->  jitter (ox, oy) of the feature point   (cellular2D.cs)
// With the thread's 4 cells in one COLUMN, Pix and everything derived from x alone (floor, residue, the three inner
// hashes, Pfx + xo) is computed once per thread and octave.  Tables: T1[Pix + c-1] = permute row offset (4 B, 128-B rows),
// T2[(T1 + Piy + j-1) mod 289] = (ox, oy) as one bank-private LDS.64 (lane l owns bytes 8l..8l+7 of a 256-B row; rows
// 289 and 290 repeat rows 0 and 1 so that j = 0..2 are immediate offsets after ONE wrap).  112 KB of shared memory.
constexpr int CT_ROWS = 291;
constexpr int C_OFF_T1 = 0, C_OFF_T2 = CT_ROWS * 128;
constexpr int CELL_SMEM = CT_ROWS * 128 + CT_ROWS * 256;
constexpr uint32_t MAGIC_SHL8 = (uint32_t)(0x4B400000ull << 8);
constexpr uint32_t WRAP8 = 289u * 256;

__device__ void build_cellular_tables() {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int r = warp; r < CT_ROWS; r += nwarps) {
        // T1 row r <-> inner hash argument Pix + oi = r - 145: byte offset of T2 row (permute + 144) [+ Piy comes in the kernel]
        *reinterpret_cast<uint32_t*>(sm + C_OFF_T1 + r * 128 + lane * 4) =
            (uint32_t)(permute_int(r - 145) + 144) * 256 + lane * 8 - MAGIC_SHL8;
        // T2 row r <-> outer hash argument r - 145: jitter of its hash, with the operations of cellular_jitter() in noise_kernels.cu
        const float p = (float)permute_int(r - 145);
        const float K = 0.142857142857f, Ko = 0.428571428571f;
        const float pk = p * K;
        const float fl = floorf(pk);
        const float ox = (pk - fl) - Ko;
        const float m7 = fmaf(-floorf(fl * (1.0f / 7.0f)), 7.0f, fl);
        const float oy = fmaf(m7, K, -Ko);
        *reinterpret_cast<float2*>(sm + C_OFF_T2 + r * 256 + lane * 8) = make_float2(ox, oy);
    }
    __syncthreads();
}
__device__ __forceinline__ float2 lds_f2(uint32_t off) { return *reinterpret_cast<const float2*>(sm + C_OFF_T2 + off); }
__device__ __forceinline__ float rectify1(float v) { return (1.0f + v) * 0.5f; }

// F1*F2 of one cell from its nine squared distances d[c][j] (min network of cellular2D.cs, as noise_kernels.cu)
struct Mins {
    float d1[3], d2[3];
};
__device__ __forceinline__ void fold_column(Mins& m, int j, float d0, float d1_, float d2_) {
    const float d1a = fminf(d0, d1_);
    float t = fmaxf(d0, d1_);
    t = fminf(t, d2_);
    m.d1[j] = fminf(d1a, t);
    m.d2[j] = fmaxf(d1a, t);
}
__device__ __forceinline__ float finish_cell(Mins& m) {
    if (!(m.d1[0] < m.d1[1])) { const float t = m.d1[0]; m.d1[0] = m.d1[1]; m.d1[1] = t; }
    if (!(m.d1[0] < m.d1[2])) { const float t = m.d1[0]; m.d1[0] = m.d1[2]; m.d1[2] = t; }
    m.d1[1] = fminf(m.d1[1], m.d2[1]);
    m.d1[2] = fminf(m.d1[2], m.d2[2]);
    m.d1[1] = fminf(m.d1[1], m.d1[2]);
    m.d1[1] = fminf(m.d1[1], m.d2[0]);
    return rectify1(sqrtf(m.d1[0])) * rectify1(sqrtf(m.d1[1]));
}

template <int PAIRS>
__global__ void __launch_bounds__(PAIR_THREADS, 1) fbm_cellular_pair_kernel(float* __restrict__ dst, FractalParams p, int wshift,
                                                                           int col_blocks, int n_items) {
    build_cellular_tables();
    const int lane = threadIdx.x & 31;
    uint32_t c1 = C_OFF_T1 + 144 * 128 + lane * 4 - MAGIC_SHL7;
    asm volatile("" : "+r"(c1));
    const int tx = threadIdx.x & ((1 << wshift) - 1), ty = threadIdx.x >> wshift;
    constexpr int CELLS = 2 * PAIRS;
    const int rows_per_item = CELLS * (PAIR_THREADS >> wshift);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cb = item % col_blocks, rg = item / col_blocks;
        const int x = (cb << wshift) + tx;
        const int r0 = rg * rows_per_item + CELLS * ty;
        const float xi = ((float)x + p.posx) / p.noise_size;
        P zi[PAIRS], t[PAIRS];
#pragma unroll
        for (int q = 0; q < PAIRS; q++) {
            zi[q].x = ((float)(p.z_first + r0 + 2 * q) + p.posz) / p.noise_size;
            zi[q].y = ((float)(p.z_first + r0 + 2 * q + 1) + p.posz) / p.noise_size;
            t[q] = bc(0.0f);
        }
        float detune = 0.0f, f = 1.0f, a = p.start_amp;
        for (int i = 0; i < p.octaves; i++) {
            // ---- everything that depends on x alone: once per thread ----
            const float Px = f * xi;
            const float flx = floorf(Px);
            const float Pfx = Px - flx;
            const float rx = fmaf(-289.0f, fmaf(flx, 1.0f / 289.0f, MAGIC) - MAGIC, flx);
            const uint32_t a1 = (__float_as_uint(rx + MAGIC) << 7) + c1;        // T1 row of Pix - 1
            uint32_t t1[3];
            float xs[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                t1[c] = lds_u32(a1 + c * 128);
                xs[c] = Pfx + (0.5f - (float)c);                                  // Pfx + xo
            }
#pragma unroll
            for (int q = 0; q < PAIRS; q++) {
                const P Py = pmul(bc(f), zi[q]);
                const P fly = pfloor(Py);
                const P Pfy = psub(Py, fly);
                const P ry = pfma(bc(-289.0f), psub(pfma(fly, bc(1.0f / 289.0f), bc(MAGIC)), bc(MAGIC)), fly);
                const P by = padd(ry, bc(MAGIC));
                uint32_t h[2][3];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t yb = __float_as_uint(e ? by.y : by.x) << 8;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const uint32_t u = yb + t1[c];                            // T2 row permute(Pix+c-1) + Piy - 1 + 145, unwrapped
                        h[e][c] = min(u, u - WRAP8);
                    }
                }
                Mins mA, mB;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const P ys = psub(Pfy, bc((float)j - 0.5f));                  // Pfy - ofj
                    float dA[3], dB[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float2 oA = lds_f2(h[0][c] + j * 256), oB = lds_f2(h[1][c] + j * 256);
                        const P dx = padd(bc(xs[c]), make_float2(oA.x, oB.x));
                        const P dy = padd(ys, make_float2(oA.y, oB.y));
                        const P d = pfma(dy, dy, pmul(dx, dx));
                        dA[c] = d.x;
                        dB[c] = d.y;
                    }
                    fold_column(mA, j, dA[0], dA[1], dA[2]);
                    fold_column(mB, j, dB[0], dB[1], dB[2]);
                }
                t[q] = pfma(bc(a), make_float2(finish_cell(mA), finish_cell(mB)), t[q]);
            }
            detune += p.detune_rate;
            f *= (p.stepdown - detune);
            a *= p.G;
        }
        if (x < p.width) {
#pragma unroll
            for (int q = 0; q < PAIRS; q++) {
                const int r = r0 + 2 * q;
                if (r < p.rows) dst[(size_t)r * p.width + x] = t[q].x / p.norm;
                if (r + 1 < p.rows) dst[(size_t)(r + 1) * p.width + x] = t[q].y / p.norm;
            }
        }
    }
}

// =====================================================================================================================
// classic Perlin (cnoise2) fBm, same two techniques.  The four lattice corners hash as permute(permute(Pix + i) + Piy + j):
// T1[Pix + i] (two adjacent rows, once per thread and octave because the thread's cells share their column) and
// T2[(T1 + Piy + j) mod 289] = the NORMALISED gradient (gx * norm, gy * norm), one bank-private LDS.64 per corner.
constexpr int PT_ROWS = 290;
constexpr int P_OFF_T1 = 0, P_OFF_T2 = PT_ROWS * 128;
constexpr int PERLIN_SMEM = PT_ROWS * 128 + PT_ROWS * 256;

__device__ void build_perlin_tables() {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int r = warp; r < PT_ROWS; r += nwarps) {
        *reinterpret_cast<uint32_t*>(sm + P_OFF_T1 + r * 128 + lane * 4) =
            (uint32_t)(permute_int(r - 144) + 144) * 256 + lane * 8 - MAGIC_SHL8;
        // gradient of hash value p with the operations of gradient_fold() / cnoise2() in noise_kernels.cu
        const float p = (float)permute_int(r - 144);
        const float fr = p * 0.024390243902439f;
        const float g = fmaf(2.0f, fr - floorf(fr), -1.0f);
        const float gy = fabsf(g) - 0.5f;
        const float gx = g - ((g + MAGIC) - MAGIC);
        const float norm = fmaf(-0.85373472095314f, fmaf(gy, gy, gx * gx), 1.79284291400159f);
        *reinterpret_cast<float2*>(sm + P_OFF_T2 + r * 256 + lane * 8) = make_float2(gx * norm, gy * norm);
    }
    __syncthreads();
}
__device__ __forceinline__ float2 lds_g2(uint32_t off) { return *reinterpret_cast<const float2*>(sm + P_OFF_T2 + off); }
__device__ __forceinline__ float fade1(float t) { return t * t * t * fmaf(t, fmaf(t, 6.0f, -15.0f), 10.0f); }
__device__ __forceinline__ P pfade(P t) { return pmul(pmul(pmul(t, t), t), pfma(t, pfma(t, bc(6.0f), bc(-15.0f)), bc(10.0f))); }

template <int PAIRS>
__global__ void __launch_bounds__(PAIR_THREADS, 1) fbm_perlin_pair_kernel(float* __restrict__ dst, FractalParams p, int wshift,
                                                                         int col_blocks, int n_items) {
    build_perlin_tables();
    const int lane = threadIdx.x & 31;
    uint32_t c1 = P_OFF_T1 + 144 * 128 + lane * 4 - MAGIC_SHL7;
    asm volatile("" : "+r"(c1));
    const int tx = threadIdx.x & ((1 << wshift) - 1), ty = threadIdx.x >> wshift;
    constexpr int CELLS = 2 * PAIRS;
    const int rows_per_item = CELLS * (PAIR_THREADS >> wshift);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cb = item % col_blocks, rg = item / col_blocks;
        const int x = (cb << wshift) + tx;
        const int r0 = rg * rows_per_item + CELLS * ty;
        const float xi = ((float)x + p.posx) / p.noise_size;
        P zi[PAIRS], t[PAIRS];
#pragma unroll
        for (int q = 0; q < PAIRS; q++) {
            zi[q].x = ((float)(p.z_first + r0 + 2 * q) + p.posz) / p.noise_size;
            zi[q].y = ((float)(p.z_first + r0 + 2 * q + 1) + p.posz) / p.noise_size;
            t[q] = bc(0.0f);
        }
        float detune = 0.0f, f = 1.0f, a = p.start_amp;
        for (int i = 0; i < p.octaves; i++) {
            // ---- x alone: once per thread ----
            const float Px = f * xi;
            const float flx = floorf(Px);
            const float pfx0 = Px - flx, pfx1 = pfx0 - 1.0f;
            const float fdx = fade1(pfx0);
            const float rx = fmaf(-289.0f, fmaf(flx, 1.0f / 289.0f, MAGIC) - MAGIC, flx);
            const uint32_t a1 = (__float_as_uint(rx + MAGIC) << 7) + c1;          // T1 row of Pix
            const uint32_t t1a = lds_u32(a1), t1b = lds_u32(a1 + 128);             // permute(Pix), permute(Pix + 1)
#pragma unroll
            for (int q = 0; q < PAIRS; q++) {
                const P Py = pmul(bc(f), zi[q]);
                const P fly = pfloor(Py);
                const P pfy0 = psub(Py, fly), pfy1 = psub(pfy0, bc(1.0f));
                const P fdy = pfade(pfy0);
                const P ry = pfma(bc(-289.0f), psub(pfma(fly, bc(1.0f / 289.0f), bc(MAGIC)), bc(MAGIC)), fly);
                const P by = padd(ry, bc(MAGIC));
                float n[4][2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t yb = __float_as_uint(e ? by.y : by.x) << 8;
                    uint32_t ua = yb + t1a, ub = yb + t1b;                         // T2 rows of (x0, y0) and (x1, y0), unwrapped
                    ua = min(ua, ua - WRAP8);
                    ub = min(ub, ub - WRAP8);
                    const float fy0 = e ? pfy0.y : pfy0.x, fy1 = e ? pfy1.y : pfy1.x;
                    const float2 g0 = lds_g2(ua), g1 = lds_g2(ub), g2 = lds_g2(ua + 256), g3 = lds_g2(ub + 256);
                    n[0][e] = fmaf(g0.y, fy0, g0.x * pfx0);                        // dot2(g * norm, (fx, fy))
                    n[1][e] = fmaf(g1.y, fy0, g1.x * pfx1);
                    n[2][e] = fmaf(g2.y, fy1, g2.x * pfx0);
                    n[3][e] = fmaf(g3.y, fy1, g3.x * pfx1);
                }
                const P n0 = make_float2(n[0][0], n[0][1]), n1 = make_float2(n[1][0], n[1][1]);
                const P n2 = make_float2(n[2][0], n[2][1]), n3 = make_float2(n[3][0], n[3][1]);
                const P l01 = pfma(bc(fdx), psub(n1, n0), n0), l23 = pfma(bc(fdx), psub(n3, n2), n2);   // math.lerp
                const P l = pfma(fdy, psub(l23, l01), l01);
                // Rectify(2.3 * l) = (1 + 2.3 l) / 2.  The sum is written fma(product, 1, 1) — with the constant first ptxas
                // folds 1*1 and then contracts the multiply into the add, which changes the rounding
                const P basis = pmul(padd(pmul(bc(2.3f), l), bc(1.0f)), bc(0.5f));
                t[q] = pfma(bc(a), basis, t[q]);
            }
            detune += p.detune_rate;
            f *= (p.stepdown - detune);
            a *= p.G;
        }
        if (x < p.width) {
#pragma unroll
            for (int q = 0; q < PAIRS; q++) {
                const int r = r0 + 2 * q;
                if (r < p.rows) dst[(size_t)r * p.width + x] = t[q].x / p.norm;
                if (r + 1 < p.rows) dst[(size_t)(r + 1) * p.width + x] = t[q].y / p.norm;
            }
        }
    }
}

// =====================================================================================================================
// psrnoise (periodic simplex with rotated gradients): PeriodicPerlin (rot 0) and RotatedSimplex (rot 0.62), the basis of
// BASELINE config C4.  Same function as psrnoise2<TYPE, true> in noise_kernels.cu, bit for bit (tests compare the two).
//
// The hash chain is  g = G[ permute( permute(xw + yw/2) + yw ) ]  per simplex corner, with (xw, yw) the corner wrapped into
// the period (1010, 102).  The INNER hash argument is a half-integer, so it is not a residue class and stays arithmetic
// (five f32x2 instructions per corner and pair); its result h1 is an exact integer in [-145, 145] and yw an exact integer
// in [0, 102), so the OUTER hash and the (cos, sin) of its value are one bank-private LDS.64:
//     T[(h1 + yw + 145) mod 289] = rotated_gradient(permute_int(h1 + yw), rot)        (289 rows x 256 B = 74 KB)
// The period wrap of the first corner uses a round-to-nearest quotient (two packed adds) instead of the scalar kernel's
// truncation (an XU instruction): the remainder is exact either way, and one conditional add of the period brings both to
// the same value.  The other two corners are the first one plus a half-integer offset with ONE wrap.
// Preconditions (host-checked, psr_pair_ok): FAST hash domain and non-negative noise coordinates, which make every wrapped
// coordinate >= -1.5 and every yw >= 0.  A cell whose first corner is not >= (1, 0) (the first lattice column) recomputes
// its six wrapped coordinates with the scalar kernel's own code.
constexpr int PSR_ROWS = 289;
constexpr float PSR_PERX = 1010.0f, PSR_PERY = 102.0f;
// The INNER hash h1 = permute(xw + yw/2) is a function of the half-integer k/2 = xw + yw/2 alone, k in [-3, 2122): a second
// table, T1[k + 3] = (h1 + 145) * 256 (the T2 row offset), takes five f32x2 instructions per corner off the FP32 pipe, which
// bounds this kernel.  Lookups are random, so the table is stored 8 times, word (k + 3) * 8 + (lane & 7): lanes of different
// residue never share a bank, the four lanes of one residue collide only when their k agree modulo 4.
constexpr int PSR_T1_ENTRIES = 2128, PSR_T1_COPIES = 8;
constexpr int PSR_OFF_T1 = PSR_ROWS * 256;
constexpr int PSR_SMEM = PSR_OFF_T1 + PSR_T1_ENTRIES * PSR_T1_COPIES * 4;
constexpr float MAGIC_HALF = 6291456.0f;                                 // 1.5 * 2^22: ulp 0.5, so k = 2 (xw + yw/2) sits in the mantissa
constexpr uint32_t MAGIC_HALF_SHL5 = (uint32_t)(0x4AC00000ull << 5);

__device__ __forceinline__ float2 psr_rotated_gradient(float p, float rot) {     // rotated_gradient() of noise_kernels.cu
    float u = fmaf(p, 0.0243902439f, rot);
    u = (u - floorf(u)) * 6.28318530718f;
    return make_float2(cosf(u), sinf(u));
}

__device__ void build_psr_table(float rot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int r = warp; r < PSR_ROWS; r += nwarps)
        *reinterpret_cast<float2*>(sm + r * 256 + lane * 8) = psr_rotated_gradient((float)permute_int(r - 145), rot);
    for (int i = threadIdx.x; i < PSR_T1_ENTRIES * PSR_T1_COPIES; i += blockDim.x) {
        // permute_centered() of noise_kernels.cu on the half-integer (k - 3) / 2, operation for operation
        const float x = (float)((i / PSR_T1_COPIES) - 3) * 0.5f;
        const float u = fmaf(34.0f, x, 1.0f) * x;
        const float q = fmaf(u, 1.0f / 289.0f, MAGIC) - MAGIC;
        const float h1 = fmaf(-289.0f, q, u);
        *reinterpret_cast<uint32_t*>(sm + PSR_OFF_T1 + i * 4) = (uint32_t)(((int)h1 + 145) * 256);
    }
    __syncthreads();
}
__device__ __forceinline__ float2 lds_psr(uint32_t off) { return *reinterpret_cast<const float2*>(sm + off); }

// fmod_period<true>() of noise_kernels.cu (the slow path keeps the scalar kernel's exact code)
__device__ __forceinline__ float psr_fmod_scalar(float p, float per, float inv_per) {
    const float a = fabsf(p);
    const float q = truncf(a * inv_per);
    float r = fmaf(-per, q, a);
    r = r < 0.0f ? r + per : r;
    r = r >= per ? r - per : r;
    return copysignf(r, p);
}
// Exact selects without predicates.  For floats a, b of which at most one is negative and the wanted one is the smaller
// NON-NEGATIVE one, the unsigned minimum of the bit patterns picks it (a negative float has the sign bit set, i.e. is huge
// as an unsigned): one integer min per element instead of a compare and a select.
__device__ __forceinline__ P umin_bits(P a, P b) {
    return make_float2(__uint_as_float(min(__float_as_uint(a.x), __float_as_uint(b.x))),
                       __uint_as_float(min(__float_as_uint(a.y), __float_as_uint(b.y))));
}
// a - b without a sign flip of b (fma(b, -1, a) is the same IEEE difference)
__device__ __forceinline__ P psubm(P a, P b) { return __ffma2_rn(b, bc(-1.0f), a); }
// r < 0 ? r + per : r   for r in (-per, per)
__device__ __forceinline__ P psr_wrap_lo(P r, float per) { return umin_bits(r, padd(r, bc(per))); }
// v >= per ? v - per : v   for v in [0, 2 per)
__device__ __forceinline__ P psr_wrap_hi(P v, float per) { return umin_bits(v, padd(v, bc(-per))); }
// v >= per ? v - per : v   for v in (-per, 2 per): a negative v stays (v - per is negative too: compare, do not bit-min)
__device__ __forceinline__ P psr_wrap_hi_signed(P v, float per) {
    const P w = padd(v, bc(-per));
    return make_float2(w.x >= 0.0f ? w.x : v.x, w.y >= 0.0f ? w.y : v.y);
}
// p mod per for p >= 0 (exact): round-to-nearest quotient, remainder in (-per, per), one wrap
__device__ __forceinline__ P psr_fmod_pos(P p, float per, float inv_per) {
    const P q = padd(pfma(p, bc(inv_per), bc(MAGIC)), bc(-MAGIC));
    return psr_wrap_lo(pfma(bc(-per), q, p), per);
}
// gradients of one corner of both cells: (gx.x, gy.x) for cell A, (gx.y, gy.y) for cell B.
//   c1 = PSR_OFF_T1 + 3 * 32 + (lane & 7) * 4 - MAGIC_HALF_SHL5,  c2 = lane * 8 - MAGIC_SHL8
__device__ __forceinline__ void psr_grad(P xw, P yw, uint32_t c1, uint32_t c2, P& gx, P& gy) {
    const P kb = pfma(bc(0.5f), yw, padd(xw, bc(MAGIC_HALF)));     // (xw + yw/2) + MAGIC_HALF, exact: k in the mantissa
    const P yb = padd(yw, bc(MAGIC));                               // yw as address bits
    const uint32_t tA = lds_u32((__float_as_uint(kb.x) << 5) + c1), tB = lds_u32((__float_as_uint(kb.y) << 5) + c1);
    uint32_t uA = (__float_as_uint(yb.x) << 8) + tA + c2, uB = (__float_as_uint(yb.y) << 8) + tB + c2;   // T2 row h1 + yw + 145
    uA = min(uA, uA - WRAP8);
    uB = min(uB, uB - WRAP8);
    const float2 gA = lds_psr(uA), gB = lds_psr(uB);
    gx = make_float2(gA.x, gB.x);
    gy = make_float2(gA.y, gB.y);
}

// psrnoise2(posx, posy) for two cells that share posx; returns TWICE the basis value Rectify(psrnoise) of both.
// Differences and selects are written in the forms that cost the fewest issue slots; each is the same exact value or the
// same single IEEE rounding as the scalar kernel's expression (commented where it is not literal).
__device__ __forceinline__ P psr_pair(float posx, P posy, uint32_t c1, uint32_t c2) {
    posy = padd(posy, bc(0.001f));
    const P ux = pfma(posy, bc(0.5f), bc(posx));
    const P i0x = pfloor(ux), i0y = pfloor(posy);
    const P f0x = psubm(ux, i0x), f0y = psubm(posy, i0y);
    // The kernel is bound by the FP32 pipe (88 f32x2 instructions per pair and octave, measured 11.7 ms at 16384^2), so
    // everything that depends on the simplex half c = f0x > f0y is a SELECT (ALU pipe), not arithmetic
    const bool cA = f0x.x > f0y.x, cB = f0x.y > f0y.y;
    const P i1x = make_float2(cA ? 1.0f : 0.0f, cB ? 1.0f : 0.0f), i1y = make_float2(cA ? 0.0f : 1.0f, cB ? 0.0f : 1.0f);
    const P p0x = pfma(i0y, bc(-0.5f), i0x);                                   // fma(-i0y, 0.5, i0x)
    const P p1x = psubm(padd(p0x, i1x), make_float2(cA ? 0.0f : 0.5f, cB ? 0.0f : 0.5f)), p1y = padd(i0y, i1y);
    const P p2x = padd(p0x, bc(0.5f)), p2y = padd(i0y, bc(1.0f));
    const P d0x = psubm(bc(posx), p0x), d0y = psubm(posy, i0y);
    const P d1x = psubm(bc(posx), p1x), d1y = psubm(posy, p1y);
    const P d2x = psubm(bc(posx), p2x), d2y = psubm(posy, p2y);
    // period wrap: first corner by exact remainder, the other two by offset and one wrap (all values are exact multiples
    // of 0.5, so any exact route to the remainder gives the scalar kernel's bits)
    P xw0 = psr_fmod_pos(p0x, PSR_PERX, 1.0f / PSR_PERX), yw0 = psr_fmod_pos(i0y, PSR_PERY, 1.0f / PSR_PERY);
    P yw2 = psr_wrap_hi(padd(yw0, bc(1.0f)), PSR_PERY);
    P yw1 = make_float2(cA ? yw0.x : yw2.x, cB ? yw0.y : yw2.y);
    P xw2 = psr_wrap_hi(padd(xw0, bc(0.5f)), PSR_PERX);
    // xw0 + (c ? 1 : -0.5) lies in [-0.5, per + 1): wrap down, then up (at most one of them changes the value)
    P xw1 = psr_wrap_lo(psr_wrap_hi_signed(padd(xw0, make_float2(cA ? 1.0f : -0.5f, cB ? 1.0f : -0.5f)), PSR_PERX), PSR_PERX);
    // first lattice column (p0x < 1; p0y >= 0 always: the host proved the coordinates non-negative): the scalar
    // kernel's general path, per cell (rare)
    const bool slowA = !(p0x.x >= 1.0f), slowB = !(p0x.y >= 1.0f);
    if (slowA || slowB) {
        const float ipx = 1.0f / PSR_PERX, ipy = 1.0f / PSR_PERY;
        if (slowA) {
            xw0.x = psr_fmod_scalar(p0x.x, PSR_PERX, ipx); yw0.x = psr_fmod_scalar(i0y.x, PSR_PERY, ipy);
            xw1.x = psr_fmod_scalar(p1x.x, PSR_PERX, ipx); yw1.x = psr_fmod_scalar(p1y.x, PSR_PERY, ipy);
            xw2.x = psr_fmod_scalar(p2x.x, PSR_PERX, ipx); yw2.x = psr_fmod_scalar(p2y.x, PSR_PERY, ipy);
        }
        if (slowB) {
            xw0.y = psr_fmod_scalar(p0x.y, PSR_PERX, ipx); yw0.y = psr_fmod_scalar(i0y.y, PSR_PERY, ipy);
            xw1.y = psr_fmod_scalar(p1x.y, PSR_PERX, ipx); yw1.y = psr_fmod_scalar(p1y.y, PSR_PERY, ipy);
            xw2.y = psr_fmod_scalar(p2x.y, PSR_PERX, ipx); yw2.y = psr_fmod_scalar(p2y.y, PSR_PERY, ipy);
        }
    }
    P g0x, g0y, g1x, g1y, g2x, g2y;
    psr_grad(xw0, yw0, c1, c2, g0x, g0y);
    psr_grad(xw1, yw1, c1, c2, g1x, g1y);
    psr_grad(xw2, yw2, c1, c2, g2x, g2y);
    const P w0 = pfma(g0y, d0y, pmul(g0x, d0x)), w1 = pfma(g1y, d1y, pmul(g1x, d1x)), w2 = pfma(g2y, d2y, pmul(g2x, d2x));
    P t0 = pmax0(psubm(bc(0.8f), pfma(d0y, d0y, pmul(d0x, d0x))));
    P t1 = pmax0(psubm(bc(0.8f), pfma(d1y, d1y, pmul(d1x, d1x))));
    P t2 = pmax0(psubm(bc(0.8f), pfma(d2y, d2y, pmul(d2x, d2x))));
    t0 = pmul(t0, t0); t0 = pmul(t0, t0);
    t1 = pmul(t1, t1); t1 = pmul(t1, t1);
    t2 = pmul(t2, t2); t2 = pmul(t2, t2);
    const P n = pmul(bc(11.0f), pfma(t2, w2, pfma(t1, w1, pmul(t0, w0))));     // 11 * dot3(t, w)
    // Rectify is (1 + v) * 0.5; the exact halving is folded into the caller's amplitude: fma(a, 0.5 x, t) == fma(0.5 a, x, t)
    return padd(n, bc(1.0f));
}

__global__ void __launch_bounds__(PAIR_THREADS, 1) fbm_psr_pair_kernel(float* __restrict__ dst, FractalParams p, float rot, int wshift,
                                                                      int col_blocks, int n_items) {
    build_psr_table(rot);
    const int lane = threadIdx.x & 31;
    uint32_t c1 = PSR_OFF_T1 + 3 * 32 + (lane & 7) * 4 - MAGIC_HALF_SHL5, c2 = lane * 8 - MAGIC_SHL8;
    asm volatile("" : "+r"(c1), "+r"(c2));
    const int tx = threadIdx.x & ((1 << wshift) - 1), ty = threadIdx.x >> wshift;
    const int rows_per_item = 4 * (PAIR_THREADS >> wshift);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cb = item % col_blocks, rg = item / col_blocks;
        const int x = (cb << wshift) + tx;
        const int r0 = rg * rows_per_item + 4 * ty;
        const float xi = ((float)x + p.posx) / p.noise_size;
        P zi[2], t[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            zi[q].x = ((float)(p.z_first + r0 + 2 * q) + p.posz) / p.noise_size;
            zi[q].y = ((float)(p.z_first + r0 + 2 * q + 1) + p.posz) / p.noise_size;
            t[q] = bc(0.0f);
        }
        float detune = 0.0f, f = 1.0f, a = p.start_amp;
        for (int i = 0; i < p.octaves; i++) {
            const float posx = f * xi;
#pragma unroll
            const float ah = a * 0.5f;            // exact (power of two): see psr_pair
#pragma unroll
            for (int q = 0; q < 2; q++) t[q] = pfma(bc(ah), psr_pair(posx, pmul(bc(f), zi[q]), c1, c2), t[q]);
            detune += p.detune_rate;
            f *= (p.stepdown - detune);
            a *= p.G;
        }
        if (x < p.width) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int r = r0 + 2 * q;
                if (r < p.rows) dst[(size_t)r * p.width + x] = t[q].x / p.norm;
                if (r + 1 < p.rows) dst[(size_t)(r + 1) * p.width + x] = t[q].y / p.norm;
            }
        }
    }
}

}  // namespace

static bool is_psr(int noise_type) { return noise_type == NZ_NOISE_PERIODIC_PERLIN || noise_type == NZ_NOISE_ROTATED_SIMPLEX; }

// the psrnoise pair kernel additionally needs non-negative noise coordinates (see its header)
bool fractal_pair_possible(int noise_type, const FractalParams& p) {
    if (!p.fast_hash || p.width < 32) return false;
    if (noise_type == NZ_NOISE_SIMPLEX || noise_type == NZ_NOISE_CELLULAR || noise_type == NZ_NOISE_PERLIN) return true;
    return is_psr(noise_type) && p.nonneg;
}

bool fractal_pair_supported(int noise_type, const FractalParams& p) {
    return fractal_pair_possible(noise_type, p) && (long long)p.width * p.rows >= (1 << 19);
}

int32_t launch_fractal_pair(float* d_dst, int noise_type, const FractalParams& p, cudaStream_t s) {
    const int sms = sm_count();
    static DeviceOnce configured;     // the attribute is per function AND per device, and sticky
    if (configured.need()) {
        NZ_CUDA(cudaFuncSetAttribute(fbm_simplex_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(fbm_cellular_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CELL_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(fbm_cellular_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CELL_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(fbm_perlin_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PERLIN_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(fbm_psr_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSR_SMEM));
        configured.mark();
    }
    int wshift = 5;
    while ((1 << wshift) < p.width && wshift < 10) wshift++;
    const int col_blocks = cdiv(p.width, 1 << wshift);
    // cells per thread: 4 (two pairs) for simplex; cellular carries nine distances per cell, NZ_CELL_PAIRS picks 1 or 2 pairs
    static const int cell_pairs = [] {
        const char* e = getenv("NZ_CELL_PAIRS");
        return e ? atoi(e) : 2;
    }();
    const int cells = noise_type == NZ_NOISE_CELLULAR ? 2 * (cell_pairs == 1 ? 1 : 2) : 4;
    const int rows_per_item = cells * (PAIR_THREADS >> wshift);
    const long long n_items = (long long)col_blocks * cdiv(p.rows, rows_per_item);
    NZ_REQUIRE(n_items < (1ll << 31), "nz_fractal: window too large");
    const int grid = n_items < sms ? (int)n_items : sms;
    if (is_psr(noise_type)) {
        const float rot = noise_type == NZ_NOISE_ROTATED_SIMPLEX ? 0.62f : 0.0f;     // Fractal.cs:184,201
        fbm_psr_pair_kernel<<<grid, PAIR_THREADS, PSR_SMEM, s>>>(d_dst, p, rot, wshift, col_blocks, (int)n_items);
    } else if (noise_type == NZ_NOISE_PERLIN) {
        fbm_perlin_pair_kernel<2><<<grid, PAIR_THREADS, PERLIN_SMEM, s>>>(d_dst, p, wshift, col_blocks, (int)n_items);
    } else if (noise_type == NZ_NOISE_CELLULAR) {
        if (cells == 2)
            fbm_cellular_pair_kernel<1><<<grid, PAIR_THREADS, CELL_SMEM, s>>>(d_dst, p, wshift, col_blocks, (int)n_items);
        else
            fbm_cellular_pair_kernel<2><<<grid, PAIR_THREADS, CELL_SMEM, s>>>(d_dst, p, wshift, col_blocks, (int)n_items);
    } else {
        fbm_simplex_pair_kernel<<<grid, PAIR_THREADS, PAIR_SMEM, s>>>(d_dst, p, wshift, col_blocks, (int)n_items);
    }
    NZ_LAUNCHED();
    return NZ_OK;
}

}  // namespace nz
