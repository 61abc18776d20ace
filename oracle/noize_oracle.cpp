/*
 * noize_oracle.cpp — CPU ORACLE for the heightmap hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file restates, in plain scalar C++, the arithmetic of xshazwar/noize-job's per-cell
 * pipeline (noise -> separable kernel filter -> flow map -> value erosion -> mesh) so the CUDA
 * kernels in noize-job_b200/csrc can be checked against it.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load the library built from it.  The
 * product (libnoize_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY STATUS: pinned at IMAGE precision, **parity unpinned** at bit level.  The reference's README screenshots
 * (docs~/0.jpg, 3.jpg, 4.jpg, 5.jpg) show rendered outputs next to the inspector panel with every parameter of the
 * run; tests/test_screenshot_pins.py runs this oracle with those parameters and rank-correlates it with the renders
 * (simplex fBm 0.93, cellular fBm 0.99, Gauss -> value erosion terrain 0.87, flow map 0.88, with controls: other
 * basis / position / size / orientation / stage order fit clearly worse).  Beyond that the reference has no tests, no
 * golden arrays and cannot be built here (Unity 2020.3 + Burst 1.5.4; no dotnet/mono/Unity in the
 * image).  The basis functions live in the un-vendored dependency com.unity.mathematics@1.2.1
 * (package.json:18), class Unity.Mathematics.noise — a C# port of the MIT-licensed Ashima Arts /
 * Stefan Gustavson "webgl-noise" GLSL library; the functions below restate that published
 * algorithm.  What IS pinned: the literal kernel tables of Filter/Kernel/KernelJob.cs:97-136 and
 * Filter/Kernel/Blur/BlurKernels.cs:59-316 (tests/golden/kernel_tables.npz, extracted from the
 * reference by tests/golden/make_golden.py), the mesh index closed form, the min-filter identity
 * and the scipy cross-checks in tests/test_oracle.py.
 *
 * Canonical evaluation order.  Burst compiles the reference with FloatMode.Fast
 * (Noise/Fractal/Fractal.cs:19), so the reference itself has no single bit pattern.  The oracle
 * fixes one: IEEE binary32, left-to-right as written in the source, with a fused multiply-add
 * exactly where this file writes fmaf() (dot products, a*b+c forms) and nowhere else — it must be
 * compiled with -ffp-contract=off.  The CUDA kernels reproduce this order op for op.
 *
 * Every function cites the reference file:line it follows (paths relative to the upstream repo).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NZREF_API extern "C" __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Unity.Mathematics.noise — common.cs (webgl-noise: mod289 / permute / taylorInvSqrt / fade)
 * ------------------------------------------------------------------------------------------ */
static inline float mod289(float x) { return fmaf(-floorf(x * (1.0f / 289.0f)), 289.0f, x); }
static inline float mod7(float x) { return fmaf(-floorf(x * (1.0f / 7.0f)), 7.0f, x); }
static inline float permute(float x) { return mod289(fmaf(34.0f, x, 1.0f) * x); }
static inline float taylorInvSqrt(float r) { return fmaf(-0.85373472095314f, r, 1.79284291400159f); }
static inline float fade(float t) { return t * t * t * fmaf(t, fmaf(t, 6.0f, -15.0f), 10.0f); }
static inline float fracf_(float x) { return x - floorf(x); }                    /* math.frac */
static inline float lerpf_(float a, float b, float t) { return fmaf(t, b - a, a); } /* math.lerp */
static inline float stepf_(float edge, float x) { return x >= edge ? 1.0f : 0.0f; } /* math.step */
static inline float dot2(float ax, float ay, float bx, float by) { return fmaf(ay, by, ax * bx); }
static inline float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaf(az, bz, fmaf(ay, by, ax * bx));
}
static inline float dot4(float ax, float ay, float az, float aw, float bx, float by, float bz, float bw) {
    return fmaf(aw, bw, fmaf(az, bz, fmaf(ay, by, ax * bx)));
}

/* noise.snoise(float2) — noise2D.cs (webgl-noise noise2D.glsl).  Call site Fractal.cs:234.
 * Returns dot(m, g) WITHOUT the final factor 130 so the getter can fold it (see basis_value case 3). */
static float snoise2_raw(float vx, float vy) {
    const float Cx = 0.211324865405187f, Cy = 0.366025403784439f;
    const float Cz = -0.577350269189626f, Cw = 0.024390243902439f;
    float s = dot2(vx, vy, Cy, Cy);
    float ix = floorf(vx + s), iy = floorf(vy + s);
    float t = dot2(ix, iy, Cx, Cx);
    float x0x = vx - ix + t, x0y = vy - iy + t;
    float i1x = x0x > x0y ? 1.0f : 0.0f, i1y = x0x > x0y ? 0.0f : 1.0f;
    float x1x = x0x + Cx - i1x, x1y = x0y + Cx - i1y;
    float x2x = x0x + Cz, x2y = x0y + Cz;
    ix = mod289(ix);
    iy = mod289(iy);
    float p[3] = {permute(permute(iy + 0.0f) + ix + 0.0f), permute(permute(iy + i1y) + ix + i1x),
                  permute(permute(iy + 1.0f) + ix + 1.0f)};
    /* max(0.5 - dot(x,x), 0): canonical contraction 0.5 - x*x - y*y -> fma(-y,y, fma(-x,x, 0.5)) */
    float m[3] = {fmaxf(fmaf(-x0y, x0y, fmaf(-x0x, x0x, 0.5f)), 0.0f), fmaxf(fmaf(-x1y, x1y, fmaf(-x1x, x1x, 0.5f)), 0.0f),
                  fmaxf(fmaf(-x2y, x2y, fmaf(-x2x, x2x, 0.5f)), 0.0f)};
    float a0[3], h[3];
    for (int k = 0; k < 3; k++) {
        m[k] = m[k] * m[k];
        m[k] = m[k] * m[k];
        float x = fmaf(2.0f, fracf_(p[k] * Cw), -1.0f);
        h[k] = fabsf(x) - 0.5f;
        float ox = floorf(x + 0.5f);
        a0[k] = x - ox;
        m[k] = m[k] * taylorInvSqrt(fmaf(h[k], h[k], a0[k] * a0[k]));
    }
    float g0 = fmaf(h[0], x0y, a0[0] * x0x);
    float g1 = fmaf(h[1], x1y, a0[1] * x1x);
    float g2 = fmaf(h[2], x2y, a0[2] * x2x);
    return dot3(m[0], m[1], m[2], g0, g1, g2);
}
/* noise.snoise(float2) = 130 * dot(m, g) */
static float snoise2(float vx, float vy) { return 130.0f * snoise2_raw(vx, vy); }

/* noise.cnoise(float2) — classicnoise2D.cs.  Call site Fractal.cs:147. */
static float cnoise2(float Px, float Py) {
    float flx = floorf(Px), fly = floorf(Py);
    float Pix[2] = {flx, flx + 1.0f}, Piy[2] = {fly, fly + 1.0f};
    float Pfx[2] = {Px - flx, (Px - flx) - 1.0f}, Pfy[2] = {Py - fly, (Py - fly) - 1.0f};
    for (int k = 0; k < 2; k++) {
        Pix[k] = mod289(Pix[k]);
        Piy[k] = mod289(Piy[k]);
    }
    /* lane order xzxz / yyww: (x0,y0) (x1,y0) (x0,y1) (x1,y1) */
    float gx[4], gy[4], fx[4], fy[4];
    for (int k = 0; k < 4; k++) {
        int kx = k & 1, ky = k >> 1;
        fx[k] = Pfx[kx];
        fy[k] = Pfy[ky];
        float i = permute(permute(Pix[kx]) + Piy[ky]);
        float g = fmaf(fracf_(i * (1.0f / 41.0f)), 2.0f, -1.0f);
        gy[k] = fabsf(g) - 0.5f;
        float tx = floorf(g + 0.5f);
        gx[k] = g - tx;
    }
    /* g00=lane0 g10=lane1 g01=lane2 g11=lane3 */
    float n[4];
    for (int k = 0; k < 4; k++) {
        float norm = taylorInvSqrt(dot2(gx[k], gy[k], gx[k], gy[k]));
        float ggx = gx[k] * norm, ggy = gy[k] * norm;
        n[k] = dot2(ggx, ggy, fx[k], fy[k]);
    }
    float fdx = fade(Pfx[0]), fdy = fade(Pfy[0]);
    float nx0 = lerpf_(n[0], n[1], fdx); /* lerp(n00, n10, fade.x) */
    float nx1 = lerpf_(n[2], n[3], fdx); /* lerp(n01, n11, fade.x) */
    return 2.3f * lerpf_(nx0, nx1, fdy);
}

/* noise.psrnoise(float2 pos, float2 per, float rot) — psrdnoise2D.cs.  Call sites Fractal.cs:184,201.
 * The Unity port wraps with math.fmod (truncating) and uses the sin/cos rgrad2 variant. */
static inline void rgrad2(float px, float py, float rot, float* gx, float* gy) {
    float u = fmaf(permute(permute(px) + py), 0.0243902439f, rot);
    u = fracf_(u) * 6.28318530718f;
    *gx = cosf(u);
    *gy = sinf(u);
}
static float psrnoise2(float posx, float posy, float perx, float pery, float rot) {
    posy += 0.001f;
    float ux = fmaf(posy, 0.5f, posx), uy = posy;
    float i0x = floorf(ux), i0y = floorf(uy);
    float f0x = ux - i0x, f0y = uy - i0y;
    float i1x = f0x > f0y ? 1.0f : 0.0f, i1y = f0x > f0y ? 0.0f : 1.0f;
    float p0x = fmaf(-i0y, 0.5f, i0x), p0y = i0y;
    float p1x = p0x + i1x - i1y * 0.5f, p1y = p0y + i1y;
    float p2x = p0x + 0.5f, p2y = p0y + 1.0f;
    float d0x = posx - p0x, d0y = posy - p0y;
    float d1x = posx - p1x, d1y = posy - p1y;
    float d2x = posx - p2x, d2y = posy - p2y;
    float xw[3] = {fmodf(p0x, perx), fmodf(p1x, perx), fmodf(p2x, perx)};
    float yw[3] = {fmodf(p0y, pery), fmodf(p1y, pery), fmodf(p2y, pery)};
    float gx[3], gy[3];
    for (int k = 0; k < 3; k++) rgrad2(fmaf(0.5f, yw[k], xw[k]), yw[k], rot, &gx[k], &gy[k]);
    float w0 = dot2(gx[0], gy[0], d0x, d0y), w1 = dot2(gx[1], gy[1], d1x, d1y), w2 = dot2(gx[2], gy[2], d2x, d2y);
    float t0 = fmaxf(0.8f - dot2(d0x, d0y, d0x, d0y), 0.0f);
    float t1 = fmaxf(0.8f - dot2(d1x, d1y, d1x, d1y), 0.0f);
    float t2 = fmaxf(0.8f - dot2(d2x, d2y, d2x, d2y), 0.0f);
    t0 = t0 * t0; t0 = t0 * t0;
    t1 = t1 * t1; t1 = t1 * t1;
    t2 = t2 * t2; t2 = t2 * t2;
    return 11.0f * dot3(t0, t1, t2, w0, w1, w2);
}

/* noise.cellular(float2) -> (F1,F2) — cellular2D.cs.  Call site Fractal.cs:269. */
static void cellular2(float Px, float Py, float* F1, float* F2) {
    const float K = 0.142857142857f, Ko = 0.428571428571f;
    float flx = floorf(Px), fly = floorf(Py);
    float Pix = mod289(flx), Piy = mod289(fly);
    float Pfx = Px - flx, Pfy = Py - fly;
    const float oi[3] = {-1.0f, 0.0f, 1.0f}, of[3] = {-0.5f, 0.5f, 1.5f};
    const float xo[3] = {0.5f, -0.5f, -1.5f};
    float d[3][3]; /* d[c] = d1,d2,d3 ; component j = x,y,z */
    for (int c = 0; c < 3; c++) {
        float pxc = permute(Pix + oi[c]);
        for (int j = 0; j < 3; j++) {
            float p = permute(pxc + Piy + oi[j]);
            float ox = fracf_(p * K) - Ko;
            float oy = fmaf(mod7(floorf(p * K)), K, -Ko);
            float dx = Pfx + xo[c] + ox; /* jitter == 1 */
            float dy = Pfy - of[j] + oy;
            d[c][j] = fmaf(dy, dy, dx * dx);
        }
    }
    float d1[3], d2[3], d1a[3];
    for (int j = 0; j < 3; j++) {
        d1a[j] = fminf(d[0][j], d[1][j]);
        d2[j] = fmaxf(d[0][j], d[1][j]);
        d2[j] = fminf(d2[j], d[2][j]);
        d1[j] = fminf(d1a[j], d2[j]);
        d2[j] = fmaxf(d1a[j], d2[j]);
    }
    if (!(d1[0] < d1[1])) { float t = d1[0]; d1[0] = d1[1]; d1[1] = t; }
    if (!(d1[0] < d1[2])) { float t = d1[0]; d1[0] = d1[2]; d1[2] = t; }
    d1[1] = fminf(d1[1], d2[1]);
    d1[2] = fminf(d1[2], d2[2]);
    d1[1] = fminf(d1[1], d1[2]);
    d1[1] = fminf(d1[1], d2[0]);
    *F1 = sqrtf(d1[0]);
    *F2 = sqrtf(d1[1]);
}

/* noise.snoise(float3) — noise3D.cs.  Call site Fractal.cs:254. */
static float snoise3(float vx, float vy, float vz) {
    const float Cx = 1.0f / 6.0f, Cy = 1.0f / 3.0f;
    float s = dot3(vx, vy, vz, Cy, Cy, Cy);
    float i[3] = {floorf(vx + s), floorf(vy + s), floorf(vz + s)};
    float t = dot3(i[0], i[1], i[2], Cx, Cx, Cx);
    float x0[3] = {vx - i[0] + t, vy - i[1] + t, vz - i[2] + t};
    /* g = step(x0.yzx, x0.xyz); l = 1 - g; i1 = min(g.xyz, l.zxy); i2 = max(g.xyz, l.zxy) */
    float g[3] = {stepf_(x0[1], x0[0]), stepf_(x0[2], x0[1]), stepf_(x0[0], x0[2])};
    float l[3] = {1.0f - g[0], 1.0f - g[1], 1.0f - g[2]};
    float i1[3] = {fminf(g[0], l[2]), fminf(g[1], l[0]), fminf(g[2], l[1])};
    float i2[3] = {fmaxf(g[0], l[2]), fmaxf(g[1], l[0]), fmaxf(g[2], l[1])};
    float x1[3], x2[3], x3[3];
    for (int k = 0; k < 3; k++) {
        x1[k] = x0[k] - i1[k] + Cx;
        x2[k] = x0[k] - i2[k] + Cy;
        x3[k] = x0[k] - 0.5f;
    }
    for (int k = 0; k < 3; k++) i[k] = mod289(i[k]);
    const float oz[4] = {0.0f, i1[2], i2[2], 1.0f}, oy[4] = {0.0f, i1[1], i2[1], 1.0f}, ox[4] = {0.0f, i1[0], i2[0], 1.0f};
    const float n_ = 0.142857142857f;
    const float nsx = n_ * 2.0f - 0.0f, nsy = n_ * 0.5f - 1.0f, nsz = n_ * 1.0f - 0.0f;
    float X[4], Y[4], H[4];
    for (int k = 0; k < 4; k++) {
        float p = permute(permute(permute(i[2] + oz[k]) + i[1] + oy[k]) + i[0] + ox[k]);
        float j = fmaf(-49.0f, floorf(p * nsz * nsz), p);
        float x_ = floorf(j * nsz);
        float y_ = floorf(fmaf(-7.0f, x_, j));
        X[k] = fmaf(x_, nsx, nsy);
        Y[k] = fmaf(y_, nsx, nsy);
        H[k] = 1.0f - fabsf(X[k]) - fabsf(Y[k]);
    }
    /* b0=(x.xy,y.xy) b1=(x.zw,y.zw); s = floor(b)*2+1; sh = -step(h,0);
       a0 = b0.xzyw + s0.xzyw*sh.xxyy ; a1 = b1.xzyw + s1.xzyw*sh.zzww
       => per corner k: ax = X[k] + (floor(X[k])*2+1)*sh[k], ay likewise with Y */
    float P[4][3];
    for (int k = 0; k < 4; k++) {
        float sh = -stepf_(H[k], 0.0f);
        float sx = fmaf(floorf(X[k]), 2.0f, 1.0f), sy = fmaf(floorf(Y[k]), 2.0f, 1.0f);
        P[k][0] = fmaf(sx, sh, X[k]);
        P[k][1] = fmaf(sy, sh, Y[k]);
        P[k][2] = H[k];
        float norm = taylorInvSqrt(dot3(P[k][0], P[k][1], P[k][2], P[k][0], P[k][1], P[k][2]));
        P[k][0] *= norm; P[k][1] *= norm; P[k][2] *= norm;
    }
    const float* xs[4] = {x0, x1, x2, x3};
    float m[4], pd[4];
    for (int k = 0; k < 4; k++) {
        m[k] = fmaxf(0.6f - dot3(xs[k][0], xs[k][1], xs[k][2], xs[k][0], xs[k][1], xs[k][2]), 0.0f);
        m[k] = m[k] * m[k];
        m[k] = m[k] * m[k];
        pd[k] = dot3(P[k][0], P[k][1], P[k][2], xs[k][0], xs[k][1], xs[k][2]);
    }
    return 42.0f * dot4(m[0], m[1], m[2], m[3], pd[0], pd[1], pd[2], pd[3]);
}

/* noise.cnoise(float3) — classicnoise3D.cs.  Call site Fractal.cs:167. */
static float cnoise3(float Px, float Py, float Pz) {
    float P[3] = {Px, Py, Pz};
    float Pi0[3], Pi1[3], Pf0[3], Pf1[3];
    for (int k = 0; k < 3; k++) {
        float fl = floorf(P[k]);
        Pi0[k] = mod289(fl);
        Pi1[k] = mod289(fl + 1.0f);
        Pf0[k] = P[k] - fl;
        Pf1[k] = Pf0[k] - 1.0f;
    }
    /* lane order: ix=(x0,x1,x0,x1) iy=(y0,y0,y1,y1); slab 0 uses z0, slab 1 uses z1 */
    float G[2][4][3];
    for (int sl = 0; sl < 2; sl++) {
        float iz = sl ? Pi1[2] : Pi0[2];
        for (int k = 0; k < 4; k++) {
            float ix = (k & 1) ? Pi1[0] : Pi0[0], iy = (k >> 1) ? Pi1[1] : Pi0[1];
            float ixy = permute(permute(ix) + iy);
            float ixyz = permute(ixy + iz);
            float gx = ixyz * (1.0f / 7.0f);
            float gy = fracf_(floorf(gx) * (1.0f / 7.0f)) - 0.5f;
            gx = fracf_(gx);
            float gz = 0.5f - fabsf(gx) - fabsf(gy);
            float sz = stepf_(gz, 0.0f);
            gx = gx - sz * (stepf_(0.0f, gx) - 0.5f);
            gy = gy - sz * (stepf_(0.0f, gy) - 0.5f);
            float norm = taylorInvSqrt(dot3(gx, gy, gz, gx, gy, gz));
            G[sl][k][0] = gx * norm; G[sl][k][1] = gy * norm; G[sl][k][2] = gz * norm;
        }
    }
    float n[2][4];
    for (int sl = 0; sl < 2; sl++)
        for (int k = 0; k < 4; k++) {
            float fx = (k & 1) ? Pf1[0] : Pf0[0], fy = (k >> 1) ? Pf1[1] : Pf0[1], fz = sl ? Pf1[2] : Pf0[2];
            n[sl][k] = dot3(G[sl][k][0], G[sl][k][1], G[sl][k][2], fx, fy, fz);
        }
    float fdx = fade(Pf0[0]), fdy = fade(Pf0[1]), fdz = fade(Pf0[2]);
    /* n_z = lerp((n000,n100,n010,n110),(n001,n101,n011,n111), fade.z) ; lanes k=0..3 */
    float nz[4];
    for (int k = 0; k < 4; k++) nz[k] = lerpf_(n[0][k], n[1][k], fdz);
    /* n_yz = lerp(n_z.xy, n_z.zw, fade.y) ; n_xyz = lerp(n_yz.x, n_yz.y, fade.x) */
    float nyz0 = lerpf_(nz[0], nz[2], fdy), nyz1 = lerpf_(nz[1], nz[3], fdy);
    return 2.2f * lerpf_(nyz0, nyz1, fdx);
}

/* ------------------------------------------------------------------------------------------
 * Noise/Fractal/Fractal.cs — basis getters (:141-278) and the fBm loop (:114-131)
 * ------------------------------------------------------------------------------------------ */
static inline float rectify(float v) { return (1.0f + v) * 0.5f; } /* (RV + v) / 2 * RV, RV = 1 */

static inline float basis_value(int type, float x, float z) {
    switch (type) {
        case 0: { /* SinGetter :210-225 */
            float vx = fmaf(0.5f, sinf(x), 0.5f), vz = fmaf(0.5f, sinf(z), 0.5f);
            return vx * vz;
        }
        case 1: return rectify(cnoise2(x, z));                              /* PerlinGetter :141-154 */
        case 2: return rectify(psrnoise2(x, z, 1010.0f, 102.0f, 0.0f));    /* PeriodicPerlinGetter :176-191 */
        case 3: /* SimplexGetter :227-241: Rectify(130*d) = (1 + 130*d)/2, canonical contraction 0.5 + 65*d */
            return fmaf(65.0f, snoise2_raw(x, z), 0.5f);
        case 4: return rectify(psrnoise2(x, z, 1010.0f, 102.0f, 0.62f));   /* RotatedSimplexGetter :193-208 */
        case 5: { /* CellularGetter :262-278 */
            float f1, f2;
            cellular2(x, z, &f1, &f2);
            return rectify(f1) * rectify(f2);
        }
        case 6:   /* PerlinGetterDomainRotated :156-174 */
        case 7: { /* SimplexGetterDomainRotated :243-260 */
            float xz = x + z;
            float s2 = xz * -0.211324865405187f;
            float xr = x + s2, zr = z + s2;
            float yr = xz * -0.577350269189626f;
            return rectify(type == 6 ? cnoise3(xr, zr, yr) : snoise3(xr, zr, yr));
        }
    }
    return 0.0f;
}

/* FractalJob.CalcFractalNormValue, Fractal.cs:31-40 */
NZREF_API float nzref_fractal_norm_value(float hurst, int32_t octaves) {
    float G = exp2f(-hurst);
    float a = 1.0f, t = 0.0f;
    for (int i = 0; i < octaves; i++) {
        t += a * 1.0f;
        a *= G;
    }
    return t;
}

/* FractalGenerator.NoiseValue / Execute, Fractal.cs:114-138; job setup :42-73 */
NZREF_API int32_t nzref_fractal(float* dst, int32_t width, int32_t rows, int32_t z_first, int32_t noise_type,
                                float hurst, float starting_amplitude, float stepdown, float detune_rate,
                                int32_t octaves, int32_t xpos, int32_t zpos, int32_t noise_size) {
    if (!dst || width <= 0 || rows <= 0 || noise_type < 0 || noise_type > 7) return -1;
    const float norm = nzref_fractal_norm_value(hurst, octaves);
    const float G = exp2f(-hurst);
    const float posx = (float)xpos, posz = (float)zpos, ns = (float)noise_size;
#pragma omp parallel for schedule(dynamic, 1)
    for (int r = 0; r < rows; r++) {
        const int z = z_first + r;
        for (int x = 0; x < width; x++) {
            float xi = ((float)x + posx) / ns;
            float zi = ((float)z + posz) / ns;
            float detune = 0.0f, f = 1.0f, a = starting_amplitude, t = 0.0f;
            for (int i = 0; i < octaves; i++) {
                float xV = f * xi, zV = f * zi;
                t = fmaf(a, basis_value(noise_type, xV, zV), t);
                detune += detune_rate;
                f *= (stepdown - detune);
                a *= G;
            }
            dst[(size_t)r * width + x] = t / norm;
        }
    }
    return 0;
}

/* single-point access to the basis functions, for known-answer tests */
NZREF_API float nzref_basis(int32_t noise_type, float x, float z) { return basis_value(noise_type, x, z); }
NZREF_API float nzref_snoise2(float x, float y) { return snoise2(x, y); }
NZREF_API float nzref_cnoise2(float x, float y) { return cnoise2(x, y); }
NZREF_API float nzref_psrnoise2(float x, float y, float perx, float pery, float rot) { return psrnoise2(x, y, perx, pery, rot); }
NZREF_API void nzref_cellular2(float x, float y, float* f) { cellular2(x, y, &f[0], &f[1]); }
NZREF_API float nzref_snoise3(float x, float y, float z) { return snoise3(x, y, z); }
NZREF_API float nzref_cnoise3(float x, float y, float z) { return cnoise3(x, y, z); }
NZREF_API float nzref_mod289(float x) { return mod289(x); }
NZREF_API float nzref_permute(float x) { return permute(x); }

/* ------------------------------------------------------------------------------------------
 * Pipeline/Tiles/TileData.cs — clamp-to-edge reads (:72-77), row-major idx (:135-138),
 * FlushWriteSlice copy-back (:16-40).  Generalised from res x res to width x rows.
 * ------------------------------------------------------------------------------------------ */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
struct Tile {
    const float* src;
    int width, rows;
    inline float get(int x, int z) const {
        x = clampi(x, 0, width - 1);
        z = clampi(z, 0, rows - 1);
        return src[(size_t)z * width + x];
    }
};
static void flush_write_slice(float* write, const float* read, size_t n) { memcpy(write, read, n * sizeof(float)); }

/* ------------------------------------------------------------------------------------------
 * Filter/Kernel — KernelOperators.cs:18-117, KernelJob.cs:18-54,165-185
 * ------------------------------------------------------------------------------------------ */
enum { PASS_SAMPLE_X, PASS_SAMPLE_Z, PASS_MIN_X, PASS_MIN_Z };

/* GenericKernelJob.ScheduleParallel (KernelJob.cs:31-53): row job src->dst then copy-back dst->src */
static void generic_kernel_job(int op, float* src, float* dst, int width, int rows, int ksize, const float* kernel,
                               float factor) {
    const int k_off = (ksize - 1) / 2;
    Tile tile = {src, width, rows};
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = 0; z < rows; z++) {
        for (int x = 0; x < width; x++) {
            float out;
            if (op == PASS_SAMPLE_X) { /* KernelSampleXOperator.ApplyKernel :31-40 */
                float total = 0.0f;
                for (int k = -k_off; k <= k_off; k++) total = fmaf(tile.get(x + k, z), kernel[k_off + k], total);
                out = total * factor;
            } else if (op == PASS_SAMPLE_Z) { /* KernelSampleZOperator.ApplyKernel :57-65 (descending k) */
                float total = 0.0f;
                for (int k = k_off; k >= -k_off; k--) total = fmaf(tile.get(x, z + k), kernel[k_off - k], total);
                out = total * factor;
            } else if (op == PASS_MIN_X) { /* KernelMinXOperator.ApplyKernel :82-91: k in [-k_off, k_off) */
                float m = 3.402823466e+38f;
                for (int k = -k_off; k < k_off; k++) { float v = tile.get(x + k, z); m = v < m ? v : m; }
                out = m;
            } else { /* KernelMinZOperator.ApplyKernel :107-117 */
                float m = 3.402823466e+38f;
                for (int k = -k_off; k < k_off; k++) { float v = tile.get(x, z + k); m = v < m ? v : m; }
                out = m;
            }
            dst[(size_t)z * width + x] = out;
        }
    }
    flush_write_slice(src, dst, (size_t)width * rows);
}

/* SeparableKernelFilter.ScheduleSeries (KernelJob.cs:165-185), iterated as KernelFilterStage.Schedule
 * does (KernelFilterStage.cs:31-43).  Result in `data`. */
NZREF_API int32_t nzref_separable(float* data, float* tmp, int32_t width, int32_t rows, int32_t ksize,
                                  const float* kx, const float* kz, float factor, int32_t iterations) {
    if (!data || !tmp || width <= 0 || rows <= 0 || ksize < 1 || !(ksize & 1) || iterations < 0) return -1;
    for (int it = 0; it < iterations; it++) {
        generic_kernel_job(PASS_SAMPLE_X, data, tmp, width, rows, ksize, kx, factor);
        generic_kernel_job(PASS_SAMPLE_Z, data, tmp, width, rows, ksize, kz, factor);
    }
    return 0;
}

/* normalised sampled Gaussian == every literal table of BlurKernels.cs:59-316 and the gauss*_s*
 * arrays of KernelJob.cs:97-105 after rounding to float (pinned by tests/golden/kernel_tables.npz) */
static void gauss_table(double sigma, int width, float* out) {
    int r = width / 2;
    double g[32], sum = 0.0;
    for (int i = -r; i <= r; i++) { g[i + r] = exp(-(double)(i * i) / (2.0 * sigma * sigma)); sum += g[i + r]; }
    for (int i = 0; i < width; i++) out[i] = (float)(g[i] / sum);
}

/* BlurHelper.limitWidth, BlurKernels.cs:30-36 */
NZREF_API int32_t nzref_limit_width(int32_t width) {
    if (width % 2 == 0) width += 1;
    if (width > 25) width = 25;
    return width < 3 ? 3 : width;
}

/* GaussianKernel.GetKernel, BlurKernels.cs:42-58 */
NZREF_API int32_t nzref_gauss_kernel(int32_t sigma, int32_t width, float* out, int32_t* width_out) {
    if (sigma < 0 || sigma > 15 || !out) return -1;
    width = nzref_limit_width(width);
    gauss_table(0.5 * (sigma + 1), width, out);
    if (width_out) *width_out = width;
    return 0;
}

/* SeparableKernelFilter.Schedule switch, KernelJob.cs:217-292 with tables :97-136 */
NZREF_API int32_t nzref_kernel_filter_table(int32_t filter, float* kx, float* kz, int32_t* ksize, float* factor) {
    static const float k101n[3] = {-1.f, 0.f, 1.f}, k121[3] = {1.f, 2.f, 1.f}, k10n1[3] = {1.f, 0.f, -1.f}, k111[3] = {1.f, 1.f, 1.f};
    int size = 3;
    float f = 1.0f;
    const float *x = k111, *z = k111;
    float g[9];
    switch (filter) {
        case 0: size = 9; gauss_table(1.0, 9, g); x = z = g; break;
        case 1: size = 7; gauss_table(1.0, 7, g); x = z = g; break;
        case 2: size = 5; gauss_table(1.0, 5, g); x = z = g; break;
        case 3: size = 3; gauss_table(1.0, 3, g); x = z = g; break;
        case 4: size = 9; gauss_table(2.0, 9, g); x = z = g; break;
        case 5: size = 7; gauss_table(2.0, 7, g); x = z = g; break;
        case 6: size = 5; gauss_table(2.0, 5, g); x = z = g; break;
        case 7: size = 3; gauss_table(2.0, 3, g); x = z = g; break;
        case 8: f = 1.0f / 3.0f; break;                 /* Smooth3 :107-108 */
        case 9: x = k101n; z = k121; break;             /* Sobel3Horizontal :110-116 */
        case 10: x = k121; z = k10n1; break;            /* Sobel3Vertical :117-122 */
        case 12: x = k10n1; z = k111; break;            /* Prewitt3Horizontal :124-130 */
        case 13: x = k111; z = k101n; break;            /* Prewitt3Vertical :131-136 */
        default: return -5;                             /* 11 = Sobel3_2D: reduce of two branches */
    }
    memcpy(kx, x, size * sizeof(float));
    memcpy(kz, z, size * sizeof(float));
    *ksize = size;
    *factor = f;
    return 0;
}

/* SeparableKernelFilter.Schedule (KernelJob.cs:217-306) x iterations (KernelFilterStage.cs:31-43).
 * Sobel3_2D follows the INTENDED semantics of ScheduleReduce (:187-215): the reference snapshots
 * `src` at schedule time (:202), a race that makes the filter "[broken]" (README.md:17); here the
 * snapshot is taken when the stage runs.  RootSumSquaresTiles: Filter/Operators/SimpleMutation.cs:148-171. */
NZREF_API int32_t nzref_kernel_filter(float* data, float* tmp, int32_t width, int32_t rows, int32_t filter,
                                      int32_t iterations) {
    if (!data || !tmp || width <= 0 || rows <= 0 || filter < 0 || filter > 13) return -1;
    float kx[9], kz[9], f;
    int ks;
    if (filter != 11) {
        nzref_kernel_filter_table(filter, kx, kz, &ks, &f);
        return nzref_separable(data, tmp, width, rows, ks, kx, kz, f, iterations);
    }
    const size_t n = (size_t)width * rows;
    float* original = (float*)malloc(n * sizeof(float));
    if (!original) return -3;
    for (int it = 0; it < iterations; it++) {
        memcpy(original, data, n * sizeof(float));
        nzref_kernel_filter_table(9, kx, kz, &ks, &f);
        nzref_separable(data, tmp, width, rows, 3, kx, kz, 1.0f, 1);
        nzref_kernel_filter_table(10, kx, kz, &ks, &f);
        nzref_separable(original, tmp, width, rows, 3, kx, kz, 1.0f, 1);
#pragma omp parallel for schedule(dynamic, 1)
        for (int z = 0; z < rows; z++)
            for (int x = 0; x < width; x++) {
                size_t i = (size_t)z * width + x;
                float a = data[i], b = original[i];
                tmp[i] = sqrtf(fmaf(b, b, a * a));
            }
        flush_write_slice(data, tmp, n);
    }
    free(original);
    return 0;
}

/* ErosionKernelJob.Schedule (KernelJob.cs:317-347) x iterations: kernelSize 3, min ops. */
NZREF_API int32_t nzref_min_erosion(float* data, float* tmp, int32_t width, int32_t rows, int32_t iterations) {
    if (!data || !tmp || width <= 0 || rows <= 0 || iterations < 0) return -1;
    for (int it = 0; it < iterations; it++) {
        generic_kernel_job(PASS_MIN_X, data, tmp, width, rows, 3, NULL, 1.0f);
        generic_kernel_job(PASS_MIN_Z, data, tmp, width, rows, 3, NULL, 1.0f);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Geologic/FlowMap — FlowMapComponents.cs:16-202, FlowMapJob.cs, Geologic/Stage/FlowMapStage.cs:124-195
 * ------------------------------------------------------------------------------------------ */
/* One FlowMapStage / ErosionStageSubtractiveFlow scratch set: water + 4 flows, each READ and WRITE
 * (FlowMapStage.cs:42-66), + tmp.  Flows start at 0 (reference: uninitialised). */
struct FlowState {
    size_t n; int width, rows;
    float *buf, *water, *water_w, *fN, *fN_w, *fS, *fS_w, *fE, *fE_w, *fW, *fW_w, *tmp;
    bool init(int w, int r) {
        width = w; rows = r; n = (size_t)w * r;
        buf = (float*)calloc(n * 11, sizeof(float));
        if (!buf) return false;
        water = buf; water_w = buf + n;
        fN = buf + 2 * n; fN_w = buf + 3 * n; fS = buf + 4 * n; fS_w = buf + 5 * n;
        fE = buf + 6 * n; fE_w = buf + 7 * n; fW = buf + 8 * n; fW_w = buf + 9 * n;
        tmp = buf + 10 * n;
        return true;
    }
    /* FillArrayJob, FlowMapComponents.cs:176-202 (FlowMapStage.cs:129) */
    void fill_water(float v) { for (size_t i = 0; i < n; i++) water[i] = v; }
    /* ComputeFlowStep.CalculateCell, FlowMapComponents.cs:20-65 */
    void outflow_step(const float* height) {
        const float TIMESTEP = 0.2f;
        Tile H = {(float*)height, width, rows}, Wt = {water, width, rows};
#pragma omp parallel for schedule(dynamic, 1)
        for (int z = 0; z < rows; z++)
            for (int x = 0; x < width; x++) {
                size_t i = (size_t)z * width + x;
                float height_0 = H.get(x, z), water_0 = Wt.get(x, z);
                float totalHt = water_0 + height_0;
                float dW = totalHt - (Wt.get(x - 1, z) + H.get(x - 1, z));
                float dE = totalHt - (Wt.get(x + 1, z) + H.get(x + 1, z));
                float dS = totalHt - (Wt.get(x, z - 1) + H.get(x, z - 1));
                float dN = totalHt - (Wt.get(x, z + 1) + H.get(x, z + 1));
                float flW = fmaxf(0.0f, fW[i] + dW), flE = fmaxf(0.0f, fE[i] + dE);
                float flS = fmaxf(0.0f, fS[i] + dS), flN = fmaxf(0.0f, fN[i] + dN);
                float sum_ = (flW + flE) + (flS + flN); /* math.csum(float4) = (x+y)+(z+w) */
                if (sum_ > 0.0f) {
                    float K = water_0 / (sum_ * TIMESTEP);
                    K = fminf(fmaxf(K, 0.0f), 1.0f);
                    fW_w[i] = flW * K; fE_w[i] = flE * K; fS_w[i] = flS * K; fN_w[i] = flN * K;
                } else {
                    fW_w[i] = 0.0f; fE_w[i] = 0.0f; fS_w[i] = 0.0f; fN_w[i] = 0.0f;
                }
            }
        /* SWAP_RWTILE x4, FlowMapJob.cs:74-77 */
        flush_write_slice(fN, fN_w, n); flush_write_slice(fS, fS_w, n);
        flush_write_slice(fE, fE_w, n); flush_write_slice(fW, fW_w, n);
    }
    /* UpdateWaterStep.CalculateCell, FlowMapComponents.cs:81-104 */
    void water_step() {
        const float TIMESTEP = 0.2f;
        Tile tN = {fN, width, rows}, tS = {fS, width, rows}, tE = {fE, width, rows}, tW = {fW, width, rows};
#pragma omp parallel for schedule(dynamic, 8)
        for (int z = 0; z < rows; z++)
            for (int x = 0; x < width; x++) {
                size_t i = (size_t)z * width + x;
                float flowOUT = ((tW.get(x, z) + tE.get(x, z)) + tS.get(x, z)) + tN.get(x, z);
                float flowIN = 0.0f;
                flowIN += tE.get(x - 1, z);
                flowIN += tW.get(x + 1, z);
                flowIN += tN.get(x, z - 1);
                flowIN += tS.get(x, z + 1);
                float ht = fmaf(flowIN - flowOUT, TIMESTEP, water[i]);
                water_w[i] = fmaxf(0.0f, ht);
            }
        flush_write_slice(water, water_w, n);
    }
    /* CreateVelocityField.CalculateCell, FlowMapComponents.cs:120-139 */
    void velocity(float* out) {
        Tile tN = {fN, width, rows}, tS = {fS, width, rows}, tE = {fE, width, rows}, tW = {fW, width, rows};
#pragma omp parallel for schedule(dynamic, 8)
        for (int z = 0; z < rows; z++)
            for (int x = 0; x < width; x++) {
                float dl = tE.get(x - 1, z) - tW.get(x, z);
                float dr = tE.get(x, z) - tW.get(x + 1, z);
                float dt = tS.get(x, z + 1) - tN.get(x, z);
                float db = tS.get(x, z) - tN.get(x, z - 1);
                float vx = (dl + dr) * 0.5f, vy = (dt + db) * 0.5f;
                out[(size_t)z * width + x] = sqrtf(fmaf(vy, vy, vx * vx));
            }
    }
    /* NormalizeMap.CalculateCell, FlowMapComponents.cs:157-165; args {min, max, max-min} FlowMapStage.cs:48-51 */
    void normalize(float* map, float norm_min, float norm_max) {
        const float a0 = norm_min, a2 = norm_max - norm_min;
#pragma omp parallel for schedule(dynamic, 8)
        for (int z = 0; z < rows; z++)
            for (int x = 0; x < width; x++) {
                size_t i = (size_t)z * width + x;
                float v = map[i];
                if (a2 < 1e-12f) v = 0.0f;
                tmp[i] = (v - a0) / a2;
            }
        flush_write_slice(map, tmp, n);
    }
};

/* FlowMapStage.ScheduleAll, Geologic/Stage/FlowMapStage.cs:124-195 */
NZREF_API int32_t nzref_flowmap(float* height, int32_t width, int32_t rows, int32_t iterations, float norm_min,
                                float norm_max) {
    if (!height || width <= 0 || rows <= 0 || iterations < 0) return -1;
    FlowState f;
    if (!f.init(width, rows)) return -3;
    f.fill_water(0.0001f);
    for (int it = 0; it < iterations; it++) {
        f.outflow_step(height);
        f.water_step();
    }
    f.velocity(height);   /* the velocity magnitude overwrites the height slice */
    f.normalize(height, norm_min, norm_max);
    free(f.buf);
    return 0;
}

/* ErosionStageSubtractiveFlow (the stage is commented-out code upstream, so this follows its text):
 * ScheduleAll (Geologic/Stage/ErosionStageSubtractiveFlow.cs:224-230) runs cycle n = 0..erosiveIterations-1 with
 * n + 1 flow iterations; ScheduleCycle (:138-222) refills the water with 1e-4 (:144), runs (outflow, water) steps on
 * flow fields that PERSIST across the cycles (allocated once, :54-76), writes the velocity magnitude into its own
 * velocityMap (:196-203), normalises it (:204-210), multiplies by erosiveFactor (ConstantMultiply, :211-217) and
 * subtracts it from the heights (SubtractTiles, :218-222).  `flowIterations` (:19-20) is never read. */
NZREF_API int32_t nzref_subtractive_flow_erosion(float* height, int32_t width, int32_t rows, int32_t erosive_iterations,
                                                 float erosive_factor, float norm_min, float norm_max) {
    if (!height || width <= 0 || rows <= 0 || erosive_iterations < 0) return -1;
    FlowState f;
    if (!f.init(width, rows)) return -3;
    float* velocity = (float*)malloc(f.n * sizeof(float));
    if (!velocity) { free(f.buf); return -3; }
    for (int n = 0; n < erosive_iterations; n++) {
        f.fill_water(0.0001f);
        for (int it = 0; it < n + 1; it++) {
            f.outflow_step(height);
            f.water_step();
        }
        f.velocity(velocity);
        f.normalize(velocity, norm_min, norm_max);
        for (size_t i = 0; i < f.n; i++) velocity[i] = velocity[i] * erosive_factor;   /* ConstantMultiply */
        for (size_t i = 0; i < f.n; i++) height[i] = height[i] - velocity[i];          /* SubtractTiles   */
    }
    free(velocity);
    free(f.buf);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Mesh — Generators/SquareGridHeightMap.cs:59-105, OvershootSquareGridHeightMap.cs:54-103,
 * Job/HeightMapMeshJob.cs:24-52 (NormalStrength = 8), Streams/PositionStream.cs:77-82,124-133
 * ------------------------------------------------------------------------------------------ */
struct Vtx { float px, py, pz, nx, ny, nz, tx, ty, tz, tw, u, v; };

NZREF_API int32_t nzref_heightmap_mesh(int32_t mesh_type, void* vertices, uint32_t* indices, int32_t R, int32_t inRes,
                                       int32_t marginPix, float Height, float TileSize, const float* heights) {
    (void)marginPix; /* only MarginScale() uses it and nothing calls MarginScale (SquareGridHeightMap.cs:41-56) */
    if (!vertices || !indices || !heights || R <= 0 || inRes <= 0 || mesh_type < 0 || mesh_type > 1) return -1;
    Vtx* vtx = (Vtx*)vertices;
    const int off = (inRes - R) / 2; /* PixOffset :33 */
    /* largest column/row the generator reads must exist (the reference would read out of bounds) */
    if (off < 0) return -1;
    if (mesh_type == 0 && R + off > inRes - 1) return -1;
    if (mesh_type == 1 && (R + 1 < R + off ? R + 1 : R + off) + off > inRes - 1) return -1;
    const float NormalStrength = 8.0f;
    const float Rf = (float)R;
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = 0; z <= R; z++) {
        int vi = (R + 1) * z, ti = 2 * R * (z - 1);
        for (int x = 0; x <= R; x++, vi++) {
            Vtx v;
            v.px = x == 0 ? -(0.5f * TileSize / Rf) : (float)x * TileSize / Rf - 0.5f;
            v.pz = (float)z * TileSize / Rf - 0.5f;
            float t, l, r, u, d;
            if (mesh_type == 0) { /* SquareGridHeightMap.getIdx :59-64, SetVertexValues :67-80 */
                auto idx = [&](int xx, int zz) {
                    xx = clampi(xx, 0, R + 1);
                    zz = clampi(zz, 0, R + 1);
                    return (size_t)(zz + off) * inRes + xx + off;
                };
                t = heights[idx(x, z)];
                l = x > 0 ? heights[idx(x - 1, z)] : t - (heights[idx(x + 1, z)] - t);
                r = x < R - 1 ? heights[idx(x + 1, z)] : t - (heights[idx(x - 1, z)] - t);
                if (z > 0) u = heights[idx(x, z - 1)];
                else { float a = heights[idx(x, z + 1)]; u = a - (t - a); }   /* InterpolateEdge(h(z+1), t) */
                if (z < R - 1) d = heights[idx(x, z + 1)];
                else { float a = heights[idx(x, z - 1)]; d = a - (t - a); }   /* InterpolateEdge(h(z-1), t) */
                v.u = (float)x / (Rf + 1.0f);
                v.v = (float)z / (Rf + 1.0f);
            } else { /* OvershootSquareGridHeightMap.getIdx :54-59, SetVertexValues :62-75 */
                auto idx = [&](int xx, int zz) {
                    xx = clampi(xx, 0 - off, R + off);
                    zz = clampi(zz, 0 - off, R + off);
                    return (size_t)(zz + off) * inRes + xx + off;
                };
                t = heights[idx(x, z)];
                l = heights[idx(x - 1, z)];
                r = heights[idx(x + 1, z)];
                u = heights[idx(x, z - 1)];
                d = heights[idx(x, z + 1)];
                v.u = (float)x / (Rf - 0.5f);
                v.v = (float)z / (Rf - 0.5f);
            }
            v.py = t * Height;
            /* t1 = (4,(r-l)/2,0), t2 = (0,(u-d)/2,4); tangent.xyz = cross(t2,t1);
               math.cross(a,b) = (a*b.yzx - a.yzx*b).yzx */
            float t1y = (r - l) / 2.0f, t2y = (u - d) / 2.0f;
            v.tx = t2y * 0.0f - 4.0f * t1y;
            v.ty = 4.0f * 4.0f - 0.0f * 0.0f;
            v.tz = 0.0f * t1y - t2y * 4.0f;
            v.tw = 0.0f;
            float nx = (l - r) / 2.0f * NormalStrength, ny = 2.0f / Height, nz = (u - d) / 2.0f * NormalStrength;
            float inv = 1.0f / sqrtf(dot3(nx, ny, nz, nx, ny, nz)); /* math.normalize = rsqrt(dot(x,x)) * x */
            v.nx = inv * nx; v.ny = inv * ny; v.nz = inv * nz;
            vtx[vi] = v;
            if (x >= 1) {
                if (z > 0) { /* SetTriangle :92-99 -> TriangleUInt32, Streams/Triangle.cs:19-28 */
                    uint32_t* tr = indices + (size_t)ti * 3;
                    tr[0] = (uint32_t)(vi - R - 2); tr[1] = (uint32_t)(vi - 1); tr[2] = (uint32_t)(vi - R - 1);
                    tr[3] = (uint32_t)(vi - R - 1); tr[4] = (uint32_t)(vi - 1); tr[5] = (uint32_t)vi;
                }
                ti += 2;
            }
        }
    }
    return 0;
}

/* MeshTileGenerator tile maths, Scripts/MeshTileGenerator.cs:166-177,197-206 */
NZREF_API int32_t nzref_tile_geometry(int32_t tileResolution, int32_t tileSize, int32_t margin, int32_t* meshResolution,
                                      int32_t* marginPix, float* meshTileSize) {
    if (tileResolution <= 0 || tileSize <= 0) return -1;
    double patchRes = (tileResolution * 1.0) / tileSize;
    int total = tileResolution + (2 * (int)(float)(margin * patchRes));
    int mv = (int)((total - tileResolution) / 2);
    float marginWS = mv * (float)((tileSize * 1.0) / tileResolution);
    if (meshResolution) *meshResolution = total;
    if (marginPix) *marginPix = mv;
    if (meshTileSize) *meshTileSize = tileSize + (2 * marginWS);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * thermal erosion and the element-wise stage set
 * ------------------------------------------------------------------------------------------ */

/* ThermalErosionFilter.rectify(float2), Filter/Kernel/Blur/ThermalErosionFilter.cs:84-98.
 * `v += increment * excess` is an a*b+c form: fmaf (canonical order, see the header). */
static inline void thermal_rectify(float* a, float* b, float maxDiff, float increment) {
    float diff = fabsf(*a - *b);
    if (diff > maxDiff) {
        float excess = diff - maxDiff;
        if (*a > *b) {
            *b = fmaf(increment, excess, *b);
            *a = fmaf(-increment, excess, *a);
        } else {
            *a = fmaf(increment, excess, *a);
            *b = fmaf(-increment, excess, *b);
        }
    }
}

/* ThermalErosionFilter.Schedule (:111-133) and Execute (:101-119): iterations x flip 0..3, one IJobFor of
 * resolution/2 - 1 row jobs per flip; idxNeighborhood / getIdx clamp (:37-50); the six pair updates of
 * rectifyNeighborhood in order xy, xz, xw, yz, yw, zw (:73-80). */
NZREF_API float nzref_thermal_max_diff(float talus_deg, float heightRatio, int32_t resolution) {
    float talus = (talus_deg / 90.0f) * 3.14159f / 2.0f;
    return (tanf(talus) * heightRatio) / (float)resolution;
}
NZREF_API int32_t nzref_thermal_erosion(float* data, int32_t resolution, float talus_deg, float incrementRatio,
                                        float meshHeightWidthRatio, int32_t iterations) {
    if (!data || resolution <= 0 || iterations < 0) return -1;
    const int res = resolution;
    const float maxDiff = nzref_thermal_max_diff(talus_deg, meshHeightWidthRatio, resolution);
    const int jobs = res / 2 - 1;
    for (int it = 0; it < iterations; it++)
        for (int flip = 0; flip < 4; flip++) {
#pragma omp parallel for schedule(dynamic, 1)
            for (int job = 0; job < jobs; job++) {
                int offset = 1;
                int z = job + 1;
                if (flip % 2 != 0) offset += 1;
                z *= 2;
                if (flip > 1) z -= 1;
                for (int x = offset; x < res - 1; x += 2) {
                    const int x1 = clampi(x + 1, 0, res - 1), z1 = clampi(z + 1, 0, res - 1);
                    float* px = &data[(size_t)z * res + x];
                    float* py = &data[(size_t)z * res + x1];
                    float* pz = &data[(size_t)z1 * res + x];
                    float* pw = &data[(size_t)z1 * res + x1];
                    float vx = *px, vy = *py, vz = *pz, vw = *pw;
                    thermal_rectify(&vx, &vy, maxDiff, incrementRatio);
                    thermal_rectify(&vx, &vz, maxDiff, incrementRatio);
                    thermal_rectify(&vx, &vw, maxDiff, incrementRatio);
                    thermal_rectify(&vy, &vz, maxDiff, incrementRatio);
                    thermal_rectify(&vy, &vw, maxDiff, incrementRatio);
                    thermal_rectify(&vz, &vw, maxDiff, incrementRatio);
                    *px = vx; *py = vy; *pz = vz; *pw = vw;
                }
            }
        }
    return 0;
}

/* ConstantJob<ConstantMultiply|ConstantBinarize>, Filter/ConstantJob.cs:16-47, SimpleMutation.cs:16-54
 * (write to tmp + SWAP_RWTILE copy-back == in place for an element-wise map) */
NZREF_API int32_t nzref_constant(float* data, int64_t n, int32_t op, float value) {
    if (!data || n < 0 || op < 0 || op > 1) return -1;
    for (int64_t i = 0; i < n; i++) data[i] = op == 0 ? data[i] * value : (data[i] >= value ? 1.0f : 0.0f);
    return 0;
}

/* ReductionJob<...>, Filter/ReductionJob.cs:16-53; operators SimpleMutation.cs:56-171; enum order
 * ReductionType, Filter/Reduce/ReduceStage.cs:12-18 */
NZREF_API int32_t nzref_reduce(float* left, const float* right, int64_t n, int32_t op) {
    if (!left || !right || n < 0 || op < 0 || op > 4) return -1;
    for (int64_t i = 0; i < n; i++) {
        const float a = left[i], b = right[i];
        float v;
        switch (op) {
            case 0: v = a - b; break;
            case 1: v = a * b; break;
            case 2: v = sqrtf(fmaf(b, b, a * a)); break;   /* sqrt((a*a) + (b*b)) */
            case 3: v = fmaxf(a, b); break;
            default: v = fminf(a, b); break;
        }
        left[i] = v;
    }
    return 0;
}

/* CurveOperator.Apply, Filter/Curve/CurveJob.cs:69-80 */
NZREF_API int32_t nzref_curve(float* data, int64_t n, const float* curve, int32_t curveSize) {
    if (!data || !curve || n < 0 || curveSize < 2) return -1;
    const float size = (float)curveSize;
    for (int64_t i = 0; i < n; i++) {
        float rect = fminf(fmaxf(data[i], 0.0f), 1.0f) * size;
        float lowerIdx = fminf(floorf(rect), size - 2.0f);
        float left = curve[(int)lowerIdx], right = curve[(int)lowerIdx + 1];
        float value = lerpf_(left, right, rect - lowerIdx);
        value = fmaxf(0.0f, value);
        value = fminf(1.0f, value);
        data[i] = value;
    }
    return 0;
}

/* CropJob.Execute, Filter/Sample/CropJob.cs:36-43 (Offset is never assigned by the reference: 0) */
NZREF_API int32_t nzref_crop(const float* input, int32_t inputResolution, float* output, int32_t outputResolution, int32_t offset) {
    if (!input || !output || inputResolution <= 0 || outputResolution <= 0) return -1;
    Tile in = {input, inputResolution, inputResolution};
    for (int z = 0; z < outputResolution; z++)
        for (int x = 0; x < outputResolution; x++) output[(size_t)z * outputResolution + x] = in.get(x + offset, z + offset);
    return 0;
}

/* GetMapRangeJob.Execute, Filter/NormalizeJob.cs:33-43 */
NZREF_API int32_t nzref_map_range(const float* map, int64_t n, float lim_min, float lim_max, float* res3) {
    if (!map || !res3 || n < 0) return -1;
    float min_ = lim_min, max_ = lim_max;
    for (int64_t i = 0; i < n; i++) {
        min_ = fminf(min_, map[i]);
        max_ = fmaxf(max_, map[i]);
    }
    res3[0] = min_; res3[1] = max_; res3[2] = max_ - min_;
    return 0;
}

/* NormalizeMap.CalculateCell, Geologic/FlowMap/FlowMapComponents.cs:157-165 via MapNormalizeValues (NormalizeJob.cs:58-92) */
NZREF_API int32_t nzref_normalize(float* data, int64_t n, const float* args3) {
    if (!data || !args3 || n < 0) return -1;
    const float a0 = args3[0], a2 = args3[2];
    for (int64_t i = 0; i < n; i++) {
        float v = data[i];
        if (a2 < 1e-12f) v = 0.0f;
        data[i] = (v - a0) / a2;
    }
    return 0;
}

NZREF_API int32_t nzref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* bench.py's reference arm sets the worker count explicitly (Unity's JobsUtility.JobWorkerCount defaults to
 * logical cores - 1): launchers such as torchrun export OMP_NUM_THREADS=1, which would otherwise decide it. */
NZREF_API int32_t nzref_set_num_threads(int32_t n) {
#ifdef _OPENMP
    if (n < 1) return -1;
    omp_set_dynamic(0);
    omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* Known-answer helper: counts integers v in [lo, hi] where the float mod289 differs from the
 * integer modulus (the float form is `x - floor(x*(1/289))*289`, which is NOT integer-mod in
 * general).  tests/test_oracle.py runs it over the whole domain the hash can reach. */
NZREF_API int64_t nzref_mod289_mismatches(int32_t lo, int32_t hi, int32_t* first_bad) {
    int64_t bad = 0;
    for (int64_t v = lo; v <= hi; v++) {
        float r = mod289((float)v);
        int64_t m = ((v % 289) + 289) % 289;
        if (r != (float)m) {
            if (!bad && first_bad) *first_bad = (int32_t)v;
            bad++;
        }
    }
    return bad;
}
/* same for mod7(floor(p*K)) vs (p div 7) mod 7 on p in [0,288] (cellular2D.cs) */
NZREF_API int32_t nzref_mod7_mismatches(void) {
    int bad = 0;
    for (int p = 0; p <= 288; p++) {
        float r = mod7(floorf((float)p * 0.142857142857f));
        if (r != (float)((p / 7) % 7)) bad++;
    }
    return bad;
}
