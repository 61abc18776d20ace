// sepwalk_kernels.cu — separable filter, T iterations per launch, entirely in registers (no shared memory).
//
// Same arithmetic and clamp semantics as filter_kernels.cu / sepfused_kernels.cu:
//   X: total = 0; for k=-r..r  total = fma(src(x+k,z), K[r+k], total); out = total*factor
//   Z: total = 0; for k=r..-r  total = fma(src(x,z+k), K[r-k], total); out = total*factor
// (KernelSampleX/ZOperator, Filter/Kernel/KernelOperators.cs:31-65; clamp Pipeline/Tiles/TileData.cs:72-77;
//  iteration loop Filter/KernelFilterStage.cs:31-43).
//
// A WARP owns a 128-column strip (lane = 4 adjacent columns) and streams down a chunk of rows.  Each of the
// T fused iterations is a pipeline stage living in registers: it receives one row per step, takes the X pass
// with its lane neighbours' values (2r warp shuffles), pushes the result into a (2r+1)-row register window and
// emits the Z pass of the window's centre row, which is the next stage's input r rows behind.  After T
// stages the row r*T behind the one just loaded is stored.  Nothing is synchronised across warps; the only
// memory traffic is one coalesced 512-byte row load and one row store per step: 8 B/cell per T iterations.
//
// Redundancy: r*T halo columns each side of the strip (112 of 128 columns useful for Gauss5 x 4) and r*T
// (+ T-1, skewed stages: see walk_body) warm-up rows per chunk.  Against the shared-memory tile kernel this removes all LDS/STS, both tile phases
// and every __syncthreads; per cell-iteration it issues 2(2r+1) FFMA + ~2r/2 SHFL.
//
// Clamp-to-edge: input rows are loaded with clamped columns; after every stage the lanes outside the grid
// take the edge lane's value (one shuffle, border warps only), and at the top/bottom of the grid a stage's
// window is filled with / keeps repeating the X pass of its first / last row.  That is exactly "an
// out-of-range neighbour reads the edge cell's current-iteration value".
#include <stdlib.h>
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int VW = 4;
constexpr int STRIP = 32 * VW;       // 128 columns per warp
constexpr int WALK_WARPS = 4;        // warps (strips) per CTA
constexpr int WALK_ZC = 128;         // rows per chunk (default; NZ_WALK_ZC overrides while profiling): 128 beat 64/96/192/256/512

template <int R>
struct TapsW {
    float k[2 * R + 1];
};

// PFR = rows of the per-warp cp.async landing ring (PFR-1 in flight); template parameter, power of two

__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem_addr, const void* gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 ld_shared_f4(unsigned smem_addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr) : "memory");
    return v;
}

// ---- bulk-copy (TMA engine) row feed, measured alternative to the per-lane cp.async feed (NZ_WALK_FEED=bulk) ----
// One elected lane issues cp.async.bulk (SASS: UBLKCP) of the warp's 512-byte row segment into the ring slot and the copy
// completes on an mbarrier of that slot; every lane then waits for the slot's phase before it reads its 16 bytes.  The
// ring becomes warp-shared, so the lanes must have finished reading a slot (__syncwarp) before lane 0 refills it.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned smem_dst, const void* gsrc, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// BORDER = false is the steady-state body for warps whose strip and chunk touch no grid border: no clamp logic.
// BULK (interior body only): rows arrive by cp.async.bulk + mbarrier instead of per-lane cp.async (see above).
// SK = 1: SKEWED stages.  Unskewed, the T stages of one step form a serial chain (stage t+1 consumes what stage t has just
// emitted: shuffle -> 2r+1 fmas -> 2r+1 fmas, T times).  Skewed, stage t+1 consumes what stage t emitted in the PREVIOUS
// step (kept in pipe[t+1]), so the T stages of a step are independent instruction streams the scheduler can interleave; a
// stage's rows arrive one step later per stage (row r0 - (R+1)t instead of r0 - R t), the chunk runs T-1 more steps, and
// every cell sees the same operations in the same order: the bits do not change.
template <int R, int T, bool SCALE, bool BORDER, int PFR, bool BULK = false, int SK = 0>
__device__ __forceinline__ void walk_body(const float* __restrict__ src, float* __restrict__ dst, int W, int H, float factor,
                                          const TapsW<R>& kx, const TapsW<R>& kz, int wx0, int zc0, int zc1, unsigned ring_base,
                                          unsigned bar_base = 0) {
    constexpr int KS = 2 * R + 1;
    constexpr int HALO = (R * T + 3) & ~3;       // multiple of 4: strips and their useful part start on a lane boundary
    constexpr int USE = STRIP - 2 * HALO;
    const int lane = threadIdx.x & 31;
    const int gx = wx0 + VW * lane;
    const bool has_left = BORDER && wx0 < 0, has_right = BORDER && wx0 + STRIP > W;
    const bool interior = !has_left && !has_right;
    const int L0 = has_left ? (-wx0) / VW : 0;                 // lane whose element 0 is grid column 0
    const int L1 = has_right ? (W - 1 - wx0) / VW : 31;        // lane whose element 3 is grid column W-1
    constexpr int LAG = R * T + SK * (T - 1);                  // steps between a row's load and its store
    int rs = max(zc0 - LAG, 0);
    rs -= rs % KS;                                             // first input row, phase 0
    // last (virtual) input row.  The clamp-free body never reads past the window: a chunk that ends at a window edge which
    // is not a grid edge (row bands, grid_edges()) simply stops there, leaving the last R*T rows unwritten — they are ghost rows
    // (skewed: the T-1 drain steps past the window's last row load nothing and push stale values through stage 0, whose
    // emissions there are ghost rows anyway)
    const int r_end = min(zc1 - 1 + LAG, BORDER ? H - 1 + LAG : H - 1 + SK * (T - 1));
    const int r_load_end = (SK && !BORDER) ? min(r_end, H - 1) : r_end;

    float win[T][KS][VW];
    float pipe[SK ? T + 1 : 1][VW];                            // skewed: pipe[t] = what stage t-1 emitted in the previous step
    if (SK) {
#pragma unroll
        for (int t = 0; t <= T; t++)
#pragma unroll
            for (int q = 0; q < VW; q++) pipe[t][q] = 0.0f;
    }
    // Memory-level parallelism.  One 512-byte row per step per warp is far too little in flight for HBM (Little's
    // law: ~12 warps x 148 SMs x 512 B / ~2 us of loaded latency ~ 0.5 TB/s).  Rows are therefore requested PFR-1 steps
    // ahead with cp.async into a per-lane private landing zone in shared memory (each lane writes and later reads its
    // own 16 bytes per row, so no cross-lane synchronisation is needed).  Border warps do the same with clamped
    // addresses — 4-byte copies where the strip overhangs the grid's first or last column; with a one-row register
    // prefetch every step of a border walk waited a full memory latency (~0.85 us: 45 us per launch).
    const unsigned ring_lane = ring_base + (unsigned)(lane * VW * sizeof(float));   // shared-space byte address
    auto request = [&](int r) {
        const unsigned slot = ring_lane + (r & (PFR - 1)) * (STRIP * 4);
        if (!BORDER) {
            cp_async16(slot, src + (size_t)r * W + gx);
        } else {
            const float* g = src + (size_t)min(r, H - 1) * W;
            if (interior) {
                cp_async16(slot, g + gx);
            } else {
#pragma unroll
                for (int q = 0; q < VW; q++) cp_async4(slot + 4 * q, g + min(max(gx + q, 0), W - 1));
            }
        }
    };
    if (BULK) {
        if (lane == 0) {
#pragma unroll 1
            for (int j = 0; j < PFR; j++) mbar_init(bar_base + j * 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll 1
            for (int j = 0; j < PFR - 1; j++)
                if (rs + j <= r_end) {
                    const unsigned slot = (rs + j) & (PFR - 1);
                    mbar_expect_tx(bar_base + slot * 8, STRIP * 4);
                    bulk_load(ring_base + slot * (STRIP * 4), src + (size_t)(rs + j) * W + wx0, STRIP * 4, bar_base + slot * 8);
                }
        }
    } else {
#pragma unroll 1
        for (int j = 0; j < PFR - 1; j++) {
            if (rs + j <= r_load_end) request(rs + j);
            cp_async_commit();
        }
    }

    for (int base = rs; base <= r_end; base += KS) {
#pragma unroll
        for (int u = 0; u < KS; u++) {
            const int r0 = base + u;
            // steady state: no guard, so the KS unrolled steps form one schedulable block (up to KS-1 surplus steps at the
            // chunk end compute rows that are never stored)
            if (!BORDER || r0 <= r_end) {
                float v[VW];
                if (BULK) {
                    const int ra = r0 + PFR - 1;                     // row requested now; its slot held row r0 - 1, read last step
                    __syncwarp();
                    if (lane == 0 && ra <= r_end) {
                        const unsigned slot = ra & (PFR - 1);
                        mbar_expect_tx(bar_base + slot * 8, STRIP * 4);
                        bulk_load(ring_base + slot * (STRIP * 4), src + (size_t)ra * W + wx0, STRIP * 4, bar_base + slot * 8);
                    }
                    // (the up to KS-1 surplus steps past r_end have no row in flight: nothing to wait for)
                    if (r0 <= r_end) mbar_wait(bar_base + (r0 & (PFR - 1)) * 8, (unsigned)(((r0 - rs) / PFR) & 1));
                    const float4 t4 = ld_shared_f4(ring_lane + (r0 & (PFR - 1)) * (STRIP * 4));
                    v[0] = t4.x; v[1] = t4.y; v[2] = t4.z; v[3] = t4.w;
                } else {
                    const int ra = r0 + PFR - 1;                     // row requested now, consumed PFR-1 steps later
                    if (ra <= r_load_end) request(ra);
                    cp_async_commit();
                    cp_async_wait<PFR - 1>();                        // the group of row r0 has landed
                    const float4 t4 = ld_shared_f4(ring_lane + (r0 & (PFR - 1)) * (STRIP * 4));
                    v[0] = t4.x; v[1] = t4.y; v[2] = t4.z; v[3] = t4.w;
                }
                float vin[VW];
#pragma unroll
                for (int q = 0; q < VW; q++) vin[q] = v[q];
                // stage order: ascending chains the stages within the step; descending (skewed) lets stage t read pipe[t] as
                // stage t-1 left it in the previous step
#pragma unroll
                for (int tt = 0; tt < T; tt++) {
                    const int t = SK ? T - 1 - tt : tt;
                    const int rt = r0 - (R + SK) * t;           // row this stage receives; phase is static:
                    const int P = ((u - (R + SK) * t) % KS + 4 * KS) % KS;
                    if (SK && t > 0) {
#pragma unroll
                        for (int q = 0; q < VW; q++) v[q] = pipe[t][q];
                    } else if (SK) {
#pragma unroll
                        for (int q = 0; q < VW; q++) v[q] = vin[q];
                    }
                    if (!BORDER || rt >= 0) {
                        float xp[VW];
                        if (!BORDER || rt <= H - 1) {
                            // X pass: a[] = columns gx-R .. gx+3+R
                            float a[VW + 2 * R];
#pragma unroll
                            for (int i = 0; i < R; i++) {
                                a[i] = __shfl_up_sync(0xffffffffu, v[VW - R + i], 1);
                                a[VW + R + i] = __shfl_down_sync(0xffffffffu, v[i], 1);
                            }
#pragma unroll
                            for (int q = 0; q < VW; q++) a[R + q] = v[q];
#pragma unroll
                            for (int q = 0; q < VW; q++) {
                                float tot = 0.0f;
#pragma unroll
                                for (int k = 0; k < KS; k++) tot = fmaf(a[q + k], kx.k[k], tot);
                                xp[q] = SCALE ? tot * factor : tot;
                            }
                        } else {
                            // below the grid: the row is a replica of the last one
#pragma unroll
                            for (int q = 0; q < VW; q++) xp[q] = win[t][(P + KS - 1) % KS][q];
                        }
                        if (BORDER && rt == 0) {
                            // above the grid: every row is a replica of row 0
#pragma unroll
                            for (int j = 0; j < KS; j++)
#pragma unroll
                                for (int q = 0; q < VW; q++) win[t][j][q] = xp[q];
                        } else {
#pragma unroll
                            for (int q = 0; q < VW; q++) win[t][P][q] = xp[q];
                        }
                        // Z pass of row rt-R: taps j = 0..2R pair K[j] with row rt-j (descending k of the reference).
                        // Packed FP32: columns (0,1) and (2,3) of the lane are one f32x2 chain each -- FFMA2 with a
                        // broadcast tap costs one issue slot for two IEEE fmas (same bits as the scalar chain).
#pragma unroll
                        for (int q = 0; q < VW; q += 2) {
                            float2 tot = make_float2(0.0f, 0.0f);
#pragma unroll
                            for (int j = 0; j < KS; j++) {
                                const int row = (P - j + 2 * KS) % KS;
                                tot = __ffma2_rn(make_float2(win[t][row][q], win[t][row][q + 1]), make_float2(kz.k[j], kz.k[j]), tot);
                            }
                            if (SCALE) tot = __fmul2_rn(tot, make_float2(factor, factor));
                            v[q] = tot.x;
                            v[q + 1] = tot.y;
                        }
                        if (has_left) {
                            const float e = __shfl_sync(0xffffffffu, v[0], L0);
                            if (lane < L0) {
#pragma unroll
                                for (int q = 0; q < VW; q++) v[q] = e;
                            }
                        }
                        if (has_right) {
                            const float e = __shfl_sync(0xffffffffu, v[VW - 1], L1);
                            if (lane > L1) {
#pragma unroll
                                for (int q = 0; q < VW; q++) v[q] = e;
                            }
                        }
                        if (SK) {
#pragma unroll
                            for (int q = 0; q < VW; q++) pipe[t + 1][q] = v[q];
                        }
                    }
                }
                if (SK) {
#pragma unroll
                    for (int q = 0; q < VW; q++) v[q] = pipe[T][q];
                }
                const int rT = r0 - LAG;
                if (rT >= zc0 && rT < zc1 && VW * lane >= HALO && VW * lane < HALO + USE && gx < W)
                    *reinterpret_cast<float4*>(dst + (size_t)rT * W + gx) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

// Interior and border warps run different BODIES: as one body the clamp logic raised every warp from 96 to 128 registers
// (Gauss5 x17 at 16384^2: 3.19 ms combined, 2.56 ms for the interior body alone, with 3 % of the warps on a border).
// Round 1 made them two launches; since the skewed interior body holds 4 CTAs per SM at up to 128 registers anyway, they
// are now two block ranges of ONE launch (sep_walk_kernel, MERGED).
struct WalkRanges {
    int s_lo, s_hi, r_lo, r_hi;   // interior launch: strips [s_lo, s_hi) x rows [r_lo, r_hi) in chunks of zc
    int ns, zcb;                  // border launch: flat (strip, chunk) items in chunks of zcb over the rest of the grid
    int n_top, n_bot, n_items;    //   all strips x [0, r_lo), all strips x [r_hi, H), border strips x [r_lo, r_hi)
    // merged launch: blocks [0, nb_first) and blocks from nb_first + n_int on take the border items (nb of them in all: either
    // all first or all last), block nb_first + by * ctas_x + bx is interior CTA (bx, by)
    int nb, nb_first, n_int, ctas_x;
};

// one (strip, chunk) item of the border list, on the clamping body
template <int R, int T, bool SCALE>
__device__ __forceinline__ void border_item(const float* __restrict__ src, float* __restrict__ dst, int W, int H, float factor,
                                            const TapsW<R>& kx, const TapsW<R>& kz, const WalkRanges& g, int item, unsigned ring_base) {
    constexpr int HALO = (R * T + 3) & ~3;
    constexpr int USE = STRIP - 2 * HALO;
    if (item >= g.n_items) return;
    int strip, zc0, zc1;
    if (item < g.n_top) {
        strip = item % g.ns;
        zc0 = (item / g.ns) * g.zcb;
        zc1 = min(zc0 + g.zcb, g.r_lo);
    } else if (item < g.n_top + g.n_bot) {
        item -= g.n_top;
        strip = item % g.ns;
        zc0 = g.r_hi + (item / g.ns) * g.zcb;
        zc1 = min(zc0 + g.zcb, H);
    } else {
        item -= g.n_top + g.n_bot;
        const int nbs = g.s_lo + (g.ns - g.s_hi);
        const int b = item % nbs;
        strip = b < g.s_lo ? b : g.s_hi + (b - g.s_lo);
        zc0 = g.r_lo + (item / nbs) * g.zcb;
        zc1 = min(zc0 + g.zcb, g.r_hi);
    }
    const int wx0 = strip * USE - HALO;
    walk_body<R, T, SCALE, true, 16>(src, dst, W, H, factor, kx, kz, wx0, zc0, zc1, ring_base);
}

// MERGED: one launch for the whole grid.  g.nb blocks (the first ones by default: see the launch) walk the border
// items on the clamping body, the others are the interior CTAs.  As two launches, even forked onto a side stream, the
// border walk did not hide under the interior launch: it cost its full ~45 us per launch at every grid size
// (tools/band_scan3.py: Gauss5 x4 on 2116 rows 135 us with the border launch, 96 us without; 16384 rows 581 / 535).
template <int R, int T, bool SCALE, int PFR, bool BULK = false, int SK = 0, bool MERGED = false>
__global__ void __launch_bounds__(WALK_WARPS * 32, ((2 * R + 1) * T > 20 || (SK && (2 * R + 1) * T > 15)) ? 4 : 5)    // window of (2R+1)*T*4 registers: 5 CTAs/SM up to 80, else 4
sep_walk_kernel(const float* __restrict__ src, float* __restrict__ dst, int W, int H, float factor, TapsW<R> kx, TapsW<R> kz,
                int zc, WalkRanges g) {
    constexpr int HALO = (R * T + 3) & ~3;
    constexpr int USE = STRIP - 2 * HALO;
    int bx = blockIdx.x, by = blockIdx.y;
    extern __shared__ __align__(16) float ring[];   // [WALK_WARPS][PFR][STRIP] (+ [WALK_WARPS][PFR] mbarriers for the bulk feed)
    const unsigned ring_base = (unsigned)__cvta_generic_to_shared(ring + (threadIdx.x >> 5) * (PFR * STRIP));
    if (MERGED) {
        int bitem = -1;
        if (bx < g.nb_first) bitem = bx;
        else if ((bx -= g.nb_first) >= g.n_int) bitem = bx - g.n_int;
        if (bitem >= 0) {
            border_item<R, T, SCALE>(src, dst, W, H, factor, kx, kz, g, bitem * WALK_WARPS + (threadIdx.x >> 5), ring_base);
            return;
        }
        by = bx / g.ctas_x;
        bx -= by * g.ctas_x;
    }
    const int strip = g.s_lo + bx * WALK_WARPS + (threadIdx.x >> 5);
    if (strip >= g.s_hi) return;                 // whole warp
    const int wx0 = strip * USE - HALO;          // grid column of this warp's column 0 (multiple of 4)
    const int zc0 = g.r_lo + by * zc, zc1 = min(zc0 + zc, g.r_hi);
    // the host chose the ranges so that the strip lies inside the grid and so does the chunk with its warm-up and drain rows
    const unsigned bar_base = (unsigned)__cvta_generic_to_shared(ring + WALK_WARPS * PFR * STRIP) + (threadIdx.x >> 5) * (PFR * 8);
    walk_body<R, T, SCALE, false, PFR, BULK, SK>(src, dst, W, H, factor, kx, kz, wx0, zc0, zc1, ring_base, bar_base);
}

template <int R, int T, bool SCALE>
__global__ void __launch_bounds__(WALK_WARPS * 32)
sep_walk_border_kernel(const float* __restrict__ src, float* __restrict__ dst, int W, int H, float factor, TapsW<R> kx,
                       TapsW<R> kz, WalkRanges g) {
    extern __shared__ __align__(16) float ring[];   // [WALK_WARPS][16][STRIP]
    const unsigned ring_base = (unsigned)__cvta_generic_to_shared(ring + (threadIdx.x >> 5) * (16 * STRIP));
    border_item<R, T, SCALE>(src, dst, W, H, factor, kx, kz, g, blockIdx.x * WALK_WARPS + (threadIdx.x >> 5), ring_base);
}

template <int R, int T>
int32_t launch_walk_rt(const float* in, float* out, int width, int rows, const float* kx, const float* kz, float factor,
                       cudaStream_t s) {
    TapsW<R> tx, tz;
    for (int i = 0; i < 2 * R + 1; i++) {
        tx.k[i] = kx[i];
        tz.k[i] = kz[i];
    }
    constexpr int HALO = (R * T + 3) & ~3;
    constexpr int USE = STRIP - 2 * HALO;
    const char* ez = getenv("NZ_WALK_ZC");
    // ranges: strip k covers grid columns [k*USE - HALO, k*USE - HALO + 128); interior strips lie inside the grid, interior
    // rows leave a band of WB rows (> R*T + 2R + 1) at the top and the bottom to the border launch
    constexpr int WB = 32;
    WalkRanges g;
    g.ns = cdiv(width, USE);
    g.s_lo = 1; g.s_hi = 0;
    for (int k = g.ns - 1; k >= 1; k--)
        if (k * USE - HALO + STRIP <= width) { g.s_hi = k + 1; break; }
    // a window edge that is not a grid edge (a row band of a larger grid: grid_edges()) needs no clamp: the interior
    // launch runs over it (what it computes within R*T rows of that edge is garbage, as the clamped values would be for the
    // band: those are ghost rows)
    const int edges = grid_edges();
    g.r_lo = (edges & 1) ? WB : 0;
    g.r_hi = (edges & 2) ? rows - WB : rows;
    if (g.s_hi <= g.s_lo || g.r_hi - g.r_lo < 32) { g.s_lo = g.s_hi = 0; g.r_lo = g.r_hi = rows; }
    g.zcb = WB;
    g.n_top = g.ns * cdiv(g.r_lo, g.zcb);
    g.n_bot = g.ns * cdiv(rows - g.r_hi, g.zcb);
    g.n_items = g.n_top + g.n_bot + (g.s_lo + (g.ns - g.s_hi)) * cdiv(g.r_hi - g.r_lo, g.zcb);
    const bool has_interior = g.s_hi > g.s_lo;
    const char* ef = getenv("NZ_WALK_FEED");      // bulk: the bulk-copy (TMA engine) row feed, kept as a MEASURED alternative (profiles/r2_walk_feed_scan.txt)
    const bool bulk = ef && ef[0] == 'b' && factor == 1.0f;
    const char* em = getenv("NZ_WALK_MERGE");     // 0: border and interior as two launches (the former form), for comparison
    const char* es = getenv("NZ_WALK_SKEW");      // 0: chained stages (the former form, which is also two launches), for comparison
    const bool chained = es && es[0] == '0';
    const bool merge = has_interior && !bulk && !chained && !(em && em[0] == '0');
    const bool skew = T > 1 && !bulk && !chained && (merge || factor == 1.0f);
    const bool forked = g.n_items > 0 && has_interior && !merge;
    const size_t ring_bytes = (size_t)WALK_WARPS * 16 * STRIP * sizeof(float);
    AuxJoinGuard side(s);                          // an early (error) return joins the side stream too
    if (g.n_items > 0 && !merge) {
        // two launches: the border launch reads the same input and writes other cells than the interior launch; it goes to
        // the side stream when there is an interior launch to run beside
        cudaStream_t bs = s;
        if (forked) {
            int32_t rc = aux_fork(s, &bs);
            if (rc != NZ_OK) return rc;
            side.armed = true;
        }
        if (factor == 1.0f)
            sep_walk_border_kernel<R, T, false><<<cdiv(g.n_items, WALK_WARPS), WALK_WARPS * 32, ring_bytes, bs>>>(in, out, width, rows, factor, tx, tz, g);
        else
            sep_walk_border_kernel<R, T, true><<<cdiv(g.n_items, WALK_WARPS), WALK_WARPS * 32, ring_bytes, bs>>>(in, out, width, rows, factor, tx, tz, g);
        NZ_LAUNCHED();
    }
    if (!has_interior) return NZ_OK;
    const int irows = g.r_hi - g.r_lo;
    const int ctas_x = cdiv(g.s_hi - g.s_lo, WALK_WARPS);
    // Rows per chunk.  A CTA walks its chunk serially and 4-5 CTAs are resident per SM, so the launch takes about
    // waves x (zc + warm-up rows) steps: take the chunk count that minimises it, within [48, 448] rows.  Measured at
    // 16384^2 (Gauss5 x17, ms): 64 2.93, 96 2.77, 128 2.68, 192 2.66, 256 2.58, 408 2.57 (two whole waves), 544 2.70.
    // (A merged launch counts its border blocks as whole CTAs: they are shorter, but a model that charges them by their
    // length picks longer chunks, which measured slower on 2645..4232-row windows: 917 us against 726 at 4232.)
    int zc = WALK_ZC;
    if (ez) {
        zc = atoi(ez);
    } else {
        const int sms = sm_count();
        const int win_regs = (2 * R + 1) * T;
        const long long slots = ((win_regs > 20 || (skew && win_regs > 15)) ? 4LL : 5LL) * sms;   // the kernel's __launch_bounds__
        const int warm = R * T + 2 * R + 1 + 3 + (skew ? T - 1 : 0);
        const long long nb = merge ? cdiv(g.n_items, WALK_WARPS) : 0;
        double best = 1e300;
        for (int n = cdiv(irows, 448); n <= irows; n++) {
            const int z = cdiv(irows, n);
            if (z < 48 && n > 1) break;
            const long long ctas = nb + (long long)ctas_x * cdiv(irows, z);
            const long long waves = (ctas + slots - 1) / slots;
            const double cost = (double)waves * (z + warm);
            if (cost < best) { best = cost; zc = z; }
        }
    }
    const size_t sm = ring_bytes;
    if (merge) {
        // Border blocks FIRST (the chunk height counts them into the waves).  Measured against border blocks last
        // (NZ_WALK_BORDER_FIRST=0), Gauss5 x17, us: 2116 rows 417 / 447, 4232 rows 731 / 804, 16384 rows 2544 / 2603.
        static const bool border_first = [] { const char* e = getenv("NZ_WALK_BORDER_FIRST"); return !(e && e[0] == '0'); }();
        g.nb = cdiv(g.n_items, WALK_WARPS);
        g.nb_first = border_first ? g.nb : 0;
        g.ctas_x = ctas_x;
        g.n_int = ctas_x * cdiv(irows, zc);
        const dim3 grid(g.nb + g.n_int);
        // (T = 1 has nothing to skew: SK = 1 is the same body there)
        if (factor == 1.0f) sep_walk_kernel<R, T, false, 16, false, 1, true><<<grid, WALK_WARPS * 32, sm, s>>>(in, out, width, rows, factor, tx, tz, zc, g);
        else sep_walk_kernel<R, T, true, 16, false, 1, true><<<grid, WALK_WARPS * 32, sm, s>>>(in, out, width, rows, factor, tx, tz, zc, g);
    } else {
        const dim3 grid(ctas_x, cdiv(irows, zc));
        if (bulk) {
            const size_t smb = sm + (size_t)WALK_WARPS * 16 * 8;
            NZ_CUDA(cudaFuncSetAttribute(sep_walk_kernel<R, T, false, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
            sep_walk_kernel<R, T, false, 16, true><<<grid, WALK_WARPS * 32, smb, s>>>(in, out, width, rows, factor, tx, tz, zc, g);
        } else if (skew) {
            sep_walk_kernel<R, T, false, 16, false, 1><<<grid, WALK_WARPS * 32, sm, s>>>(in, out, width, rows, factor, tx, tz, zc, g);
        } else if (factor == 1.0f) {
            sep_walk_kernel<R, T, false, 16><<<grid, WALK_WARPS * 32, sm, s>>>(in, out, width, rows, factor, tx, tz, zc, g);
        } else {
            sep_walk_kernel<R, T, true, 16><<<grid, WALK_WARPS * 32, sm, s>>>(in, out, width, rows, factor, tx, tz, zc, g);
        }
    }
    NZ_LAUNCHED();
    if (forked) return side.join();
    return NZ_OK;
}

template <int R>
int32_t launch_walk_r(float* d_data, float* d_tmp, int width, int rows, const float* kx, const float* kz, float factor,
                      int iterations, float** d_result, cudaStream_t s) {
    // register windows: T*(2R+1)*4 floats per lane.  NZ_WALK_TMAX=5 (R <= 2 only) tries 5 stages per launch: 17 iterations in
    // 4 launches (5,4,4,4) instead of 5 (4,4,3,3,3) — measured, see profiles/r2_walk_tmax_scan.txt
    int TMAX = R <= 2 ? 4 : 2;
    if (R <= 2) {
        const char* et = getenv("NZ_WALK_TMAX");
        if (et && atoi(et) == 5) TMAX = 5;
    }
    const int launches = (iterations + TMAX - 1) / TMAX;
    float *cur = d_data, *other = d_tmp;
    int left = iterations;
    for (int l = 0; l < launches; l++) {
        const int T = (left + (launches - l) - 1) / (launches - l);
        int32_t rc;
        switch (T) {
            case 1: rc = launch_walk_rt<R, 1>(cur, other, width, rows, kx, kz, factor, s); break;
            case 2: rc = launch_walk_rt<R, 2>(cur, other, width, rows, kx, kz, factor, s); break;
            // T >= 3 is only reached with R <= 2 (TMAX above): for R > 2 the template argument names an instantiation that
            // exists anyway, so the deep register windows of (R = 3, 4; T = 3, 4) are not compiled at all
            case 3: rc = launch_walk_rt<(R <= 2 ? R : 1), 3>(cur, other, width, rows, kx, kz, factor, s); break;
            case 5: rc = launch_walk_rt<(R <= 2 ? R : 1), 5>(cur, other, width, rows, kx, kz, factor, s); break;
            default: rc = launch_walk_rt<(R <= 2 ? R : 1), 4>(cur, other, width, rows, kx, kz, factor, s); break;
        }
        if (rc != NZ_OK) return rc;
        left -= T;
        float* t = cur; cur = other; other = t;
    }
    if (d_result) {
        *d_result = cur;
    } else if (cur != d_data) {
        NZ_CUDA(cudaMemcpyAsync(d_data, cur, (size_t)width * rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return NZ_OK;
}

}  // namespace

bool separable_walk_supported(int width, int ksize, const void* a, const void* b) {
    return ksize >= 3 && ksize <= 9 && (ksize & 1) && (width & 3) == 0 && (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

int32_t launch_separable_walk(float* d_data, float* d_tmp, int width, int rows, int ksize, const float* kx,
                              const float* kz, float factor, int iterations, float** d_result, cudaStream_t s) {
    switch (ksize) {
        case 3: return launch_walk_r<1>(d_data, d_tmp, width, rows, kx, kz, factor, iterations, d_result, s);
        case 5: return launch_walk_r<2>(d_data, d_tmp, width, rows, kx, kz, factor, iterations, d_result, s);
        case 7: return launch_walk_r<3>(d_data, d_tmp, width, rows, kx, kz, factor, iterations, d_result, s);
        case 9: return launch_walk_r<4>(d_data, d_tmp, width, rows, kx, kz, factor, iterations, d_result, s);
    }
    set_error("separable_walk: unsupported ksize %d", ksize);
    return NZ_E_UNSUPPORTED;
}

}  // namespace nz
