// sde_core.cuh - device-side building blocks of the batched SDE integrator (sm_100a).
//
// Self-contained (no host / libc headers) so that the SAME text is compiled by nvcc at build
// time and by NVRTC at run time for traced user problems.
//
// Replaces, for the hot path of the reference (PyChastic):
//   * jax.random.{PRNGKey,split,normal}  (reference pychastic/sde_solver.py:223,292,299 and
//     pychastic/vectorized_I_generation.py:267-349) -> in-register threefry2x32 + Giles erf_inv
//   * get_wiener_integrals / make_C_mat / make_D_mat
//     (reference pychastic/vectorized_I_generation.py:22-49,135-236,243-473) -> WienerGen<>
//
// Compile-time configuration (macros, defined by the including translation unit):
//   SDE_M       number of noise terms
//   SDE_SCHEME  0 euler, 1 milstein, 2 wagner_platen
//   SDE_P       Fourier series truncation (reference always uses 10, sde_solver.py:271)
//   SDE_F64     1: double (64 random bits per draw), 0: float (32 random bits per draw)
//   SDE_LAYOUT  0: original threefry counter layout (jax < 0.5), 1: partitionable (jax >= 0.5)
#pragma once

#ifndef SDE_P
#define SDE_P 10
#endif
#ifndef SDE_LAYOUT
#define SDE_LAYOUT 0
#endif

typedef unsigned int u32;
typedef unsigned long long u64;

#if SDE_F64
typedef double real;
#define RC(x) x
#else
typedef float real;
#define RC(x) x##f
#endif

#define SDE_PI 3.14159265358979323846
#define SDE_EULER 0
#define SDE_MILSTEIN 1
#define SDE_WAGNER_PLATEN 2

namespace sdeb {

struct Key {
    u32 a, b;
};

// ----------------------------------------------------------------------------------------
// threefry2x32, 20 rounds (Random123; jax/_src/prng.py::threefry2x32)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ u32 rotl32(u32 x, int r) { return __funnelshift_l(x, x, r); }

__device__ __forceinline__ Key threefry2x32(Key k, u32 x0, u32 x1) {
    const u32 ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
    x0 += ks0;
    x1 += ks1;
#define TF_ROUND(r) x0 += x1; x1 = rotl32(x1, r); x1 ^= x0;
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks1; x1 += ks2 + 1u;
    TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
    x0 += ks2; x1 += ks0 + 2u;
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks0; x1 += ks1 + 3u;
    TF_ROUND(17) TF_ROUND(29) TF_ROUND(16) TF_ROUND(24)
    x0 += ks1; x1 += ks2 + 4u;
    TF_ROUND(13) TF_ROUND(15) TF_ROUND(26) TF_ROUND(6)
    x0 += ks2; x1 += ks0 + 5u;
#undef TF_ROUND
    Key out;
    out.a = x0;
    out.b = x1;
    return out;
}

// jax.random.split(key, n)[i].  Original layout: the 2n outputs of
// threefry_2x32(key, arange(2n)) are [word0(block(k, n+k)) k<n | word1(block(k-n, k)) k>=n],
// key i = outputs (2i, 2i+1).  Partitionable: both words of block (hi32 i, lo32 i).
__device__ __forceinline__ u32 split_word(Key key, u64 n, u64 k) {
    if (k < n) return threefry2x32(key, (u32)k, (u32)(n + k)).a;
    return threefry2x32(key, (u32)(k - n), (u32)k).b;
}

__device__ __forceinline__ Key split_key(Key key, u64 n, u64 i) {
#if SDE_LAYOUT == 0
    Key out;
    out.a = split_word(key, n, 2 * i);
    out.b = split_word(key, n, 2 * i + 1);
    return out;
#else
    return threefry2x32(key, (u32)(i >> 32), (u32)i);
#endif
}

// ----------------------------------------------------------------------------------------
// random bits for element e of a draw of `size` elements (row-major flat index)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ u64 random_bits64(Key key, u32 e, u32 size) {
#if SDE_LAYOUT == 0
    Key y = threefry2x32(key, e, size + e);
#else
    Key y = threefry2x32(key, 0u, e);
#endif
    return ((u64)y.a << 32) | (u64)y.b;
}

__device__ __forceinline__ u32 random_bits32(Key key, u32 e, u32 size) {
#if SDE_LAYOUT == 0
    const u32 h = (size + 1u) >> 1;
    if (e < h) {
        u32 c1 = h + e;
        if (c1 >= size) c1 = 0u;  // odd size: the pad counter
        return threefry2x32(key, e, c1).a;
    }
    return threefry2x32(key, e - h, e).b;
#else
    Key y = threefry2x32(key, 0u, e);
    return y.a ^ y.b;
#endif
}

// ----------------------------------------------------------------------------------------
// erf_inv: XLA's expansion (Giles' polynomials), w = -log1p(-x*x)
// ----------------------------------------------------------------------------------------
// Coefficients live in constant memory: each Horner step is then ONE FMA with a c[bank][offset]
// operand instead of two immediate moves + FMA, and the unrolled polynomial stays small.
static __constant__ float ERFINV32_LT5[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f,
                                      0.00021858087f,  -0.00125372503f, -0.00417768164f, 0.246640727f,
                                      1.50140941f};
static __constant__ float ERFINV32_GE5[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f,
                                      0.00573950773f,   -0.0076224613f,  0.00943887047f, 1.00167406f,
                                      2.83297682f};
static __constant__ double ERFINV64_LT6_25[23] = {
    -3.6444120640178196996e-21, -1.685059138182016589e-19,  1.2858480715256400167e-18, 1.115787767802518096e-17,
    -1.333171662854620906e-16,  2.0972767875968561637e-17,  6.6376381343583238325e-15, -4.0545662729752068639e-14,
    -8.1519341976054721522e-14, 2.6335093153082322977e-12,  -1.2975133253453532498e-11, -5.4154120542946279317e-11,
    1.051212273321532285e-09,   -4.1126339803469836976e-09, -2.9070369957882005086e-08, 4.2347877827932403518e-07,
    -1.3654692000834678645e-06, -1.3882523362786468719e-05, 0.0001867342080340571352,  -0.00074070253416626697512,
    -0.0060336708714301490533,  0.24015818242558961693,     1.6536545626831027356};
static __constant__ double ERFINV64_LT16[19] = {
    2.2137376921775787049e-09,  9.0756561938885390979e-08,  -2.7517406297064545428e-07, 1.8239629214389227755e-08,
    1.5027403968909827627e-06,  -4.013867526981545969e-06,  2.9234449089955446044e-06,  1.2475304481671778723e-05,
    -4.7318229009055733981e-05, 6.8284851459573175448e-05,  2.4031110387097893999e-05,  -0.0003550375203628474796,
    0.00095328937973738049703,  -0.0016882755560235047313,  0.0024914420961078508066,   -0.0037512085075692412107,
    0.005370914553590063617,    1.0052589676941592334,      3.0838856104922207635};
static __constant__ double ERFINV64_GE16[17] = {
    -2.7109920616438573243e-11, -2.5556418169965252055e-10, 1.5076572693500548083e-09,  -3.7894654401267369937e-09,
    7.6157012080783393804e-09,  -1.4960026627149240478e-08, 2.9147953450901080826e-08,  -6.7711997758452339498e-08,
    2.2900482228026654717e-07,  -9.9298272942317002539e-07, 4.5260625972231537039e-06,  -1.9681778105531670567e-05,
    7.5995277030017761139e-05,  -0.00021503011930044477347, -0.00013871931833623122026, 1.0103004648645343977,
    4.8499064014085844221};

// ln2 split so that k * ln2_hi is exact for every exponent k that can occur (21 trailing zero bits)
static __constant__ double LOGC[2] = {6.93147180369123816490e-01, 1.90821492927058770002e-10};

// 64-entry table for log(m), m in [1, 2): entry i = (1/c_i rounded to double, -log of that double) with
// c_i = 1 + (i + 1/2)/64, generated with 60-digit arithmetic.  Global memory + read-only cache: the indices
// are per-lane, which the constant cache would serialise.
static __device__ const double2 SDE_LOGTAB[64] = {
    {0.9922480620155039, 0.007782140442054963}, {0.9770992366412213, 0.023167059281534418},
    {0.9624060150375939, 0.03831886430213666}, {0.9481481481481482, 0.05324451451881224},
    {0.9343065693430657, 0.06795066190850778}, {0.920863309352518, 0.08244366921107454},
    {0.9078014184397163, 0.09672962645855114}, {0.8951048951048951, 0.11081436634029011},
    {0.8827586206896552, 0.12470347850095725}, {0.8707482993197279, 0.1384023228591192},
    {0.8590604026845637, 0.151916042025842}, {0.847682119205298, 0.16524957289530717},
    {0.8366013071895425, 0.17840765747281825}, {0.8258064516129032, 0.19139485299962947},
    {0.8152866242038217, 0.20421554142869083}, {0.8050314465408805, 0.2168739383006143},
    {0.7950310559006211, 0.2293741010648459}, {0.7852760736196319, 0.24171993688714513},
    {0.7757575757575758, 0.25391520998096345}, {0.7664670658682635, 0.2659635484971379},
    {0.757396449704142, 0.2778684510034563}, {0.7485380116959064, 0.2896332925830427},
    {0.7398843930635838, 0.30126133057816185}, {0.7314285714285714, 0.3127557100038969},
    {0.7231638418079096, 0.324119468654212}, {0.7150837988826816, 0.3353555419211378},
    {0.7071823204419889, 0.3464667673462086}, {0.6994535519125683, 0.3574558889218038},
    {0.6918918918918919, 0.36832556115870757}, {0.6844919786096256, 0.3790783529349695},
    {0.6772486772486772, 0.38971675114002524}, {0.6701570680628273, 0.40024316412701266},
    {0.6632124352331606, 0.4106599249852683}, {0.6564102564102564, 0.42096929464412963},
    {0.649746192893401, 0.43117346481837143}, {0.6432160804020101, 0.4412745608048752},
    {0.6368159203980099, 0.4512746441394586}, {0.6305418719211823, 0.46117571512217015},
    {0.624390243902439, 0.470979715218791}, {0.6183574879227053, 0.48068852934575196},
    {0.6124401913875598, 0.4903039880451939}, {0.6066350710900474, 0.49982786955644926},
    {0.6009389671361502, 0.5092619017898079}, {0.5953488372093023, 0.5186077642080457},
    {0.5898617511520737, 0.5278670896208424}, {0.5844748858447488, 0.5370414658968837},
    {0.579185520361991, 0.5461324375981356}, {0.5739910313901345, 0.5551415075405016},
    {0.5688888888888889, 0.564070138284803}, {0.5638766519823789, 0.5729197535617854},
    {0.5589519650655022, 0.5816917396346225}, {0.5541125541125541, 0.5903874466021763},
    {0.5493562231759657, 0.5990081896460834}, {0.5446808510638298, 0.6075552502245418},
    {0.540084388185654, 0.616029877215514}, {0.5355648535564853, 0.6244332880118936},
    {0.5311203319502075, 0.6327666695710378}, {0.5267489711934157, 0.6410311794209312},
    {0.5224489795918368, 0.6492279466251097}, {0.5182186234817814, 0.65735807270836},
    {0.5140562248995983, 0.6654226325450905}, {0.5099601593625498, 0.6734226752121667},
    {0.5059288537549407, 0.6813592248079031}, {0.5019607843137255, 0.689233281238809},
};

static __constant__ double SDE_LOGPOLY[6] = {1.0 / 7.0, 0.2, 1.0 / 3.0, 1.4142135623730951, -1.0 / 6.0,
                                            -0.99999999999999988898 /* nextafter(-1, 0) */};

#ifndef SDE_LOGTAB_IN_SMEM
#define SDE_LOGTAB_IN_SMEM 0           // kernels that set this call sde_fill_logtab() before the first draw
#endif
#if SDE_LOGTAB_IN_SMEM
__shared__ double2 sde_logtab_smem[64];
// all threads of the CTA, before any normal is drawn (contains a __syncthreads)
__device__ __forceinline__ void sde_fill_logtab() {
    if (threadIdx.x < 64) sde_logtab_smem[threadIdx.x] = SDE_LOGTAB[threadIdx.x];
    __syncthreads();
}
#endif

// w = -log(1 - x^2) = -log1p(-x*x) for |x| < 1, branch-free.
//
// CUDA's log1p() takes one of two ~110-instruction paths depending on the argument, and the lanes of a warp
// split evenly between them, so every normal paid for both.  Here t = 1 - x^2 = 2^k m is split with integer
// ops on the exponent field, m is reduced with one table entry (r = m / c_i - 1 as a single FMA, |r| < 2^-7)
// and log(1 + r) is a degree-7 polynomial:  log t = k ln2 + log c_i + log1p(r).
// Absolute error ~1e-16 for w <= 6.25 and relative error ~1e-16 beyond, which is all erf_inv's polynomial in w
// can see (the result is as accurate as one obtained from log1p's correctly rounded w).
__device__ __forceinline__ double neg_log_one_minus_sq(double x) {
    // x*x is rounded BEFORE the subtraction, exactly as in XLA's `-log1p(-x * x)`: near |x| = 1 the
    // result is ill-conditioned in x*x, so a fused 1 - x*x would change the far-tail normals.
    const double t = 1.0 - __dmul_rn(x, x);                 // in (0, 1], always a normal number
    const int hi = __double2hiint(t);
    const int lo = __double2loint(t);
    const double dk = (double)((hi >> 20) - 1023);
#if SDE_LOGTAB_IN_SMEM
    const double2 e = sde_logtab_smem[(hi >> 14) & 63];     // 32-bit shared address: two integer ops fewer per draw
#else
    const double2 e = __ldg(&SDE_LOGTAB[(hi >> 14) & 63]);
#endif
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double r = fma(m, e.x, -1.0);
    // literals whose low word is non-zero come from the constant bank (one operand of the FMA) instead of
    // being rebuilt with two moves per use
    double q = fma(r, SDE_LOGPOLY[0], -1.0 / 6.0);
    q = fma(r, q, SDE_LOGPOLY[1]);
    q = fma(r, q, -0.25);
    q = fma(r, q, SDE_LOGPOLY[2]);
    q = fma(r, q, -0.5);
    const double p = fma(r * r, q, r);                      // log1p(r)
    // w = -(k ln2_hi + (log c + (p + k ln2_lo)))
    return fma(-dk, LOGC[0], -(e.y + fma(dk, LOGC[1], p)));
}

// erf_inv is split into the branch-free central polynomial (evaluated for a whole group of draws with
// the Horner chains interleaved) and a rarely taken tail fix-up (probability ~1e-3 per draw in double).
__device__ __forceinline__ float erfinv_w(float x) { return -log1pf(-x * x); }
__device__ __forceinline__ double erfinv_w(double x) { return neg_log_one_minus_sq(x); }

__device__ __forceinline__ bool erfinv_is_central(float w) { return w < 5.0f; }
__device__ __forceinline__ bool erfinv_is_central(double w) { return w < 6.25; }

#ifndef SDE_TAIL_ATTR
#define SDE_TAIL_ATTR __noinline__     // rarely taken: keep it out of the unrolled groups (override to __forceinline__)
#endif
static __device__ SDE_TAIL_ATTR float erfinv_tail(float x, float w) {
    w = sqrtf(w) - 3.0f;
    float p = ERFINV32_GE5[0];
#pragma unroll 1
    for (int i = 1; i < 9; ++i) p = fmaf(p, w, ERFINV32_GE5[i]);
    return (fabsf(x) == 1.0f) ? x * __int_as_float(0x7f800000) : p * x;
}

static __device__ SDE_TAIL_ATTR double erfinv_tail(double x, double w) {
    double p;
    if (w < 16.0) {
        w = sqrt(w) - 3.25;
        p = ERFINV64_LT16[0];
#pragma unroll 1
        for (int i = 1; i < 19; ++i) p = fma(p, w, ERFINV64_LT16[i]);
    } else {
        w = sqrt(w) - 5.0;
        p = ERFINV64_GE16[0];
#pragma unroll 1
        for (int i = 1; i < 17; ++i) p = fma(p, w, ERFINV64_GE16[i]);
    }
    return (fabs(x) == 1.0) ? x * __longlong_as_double(0x7ff0000000000000LL) : p * x;
}

// jax.random.uniform(key, shape, dtype, nextafter(-1,0), 1) for one element, then
// jax.random.normal = sqrt(2) * erf_inv(u)
__device__ __forceinline__ double uniform_pm1(u64 bits) {
    double f = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ULL)) - 1.0;
    // jax clamps with max(lo, .) against rounding below lo; here hi - lo rounds to exactly 2, f * 2 is exact and
    // f >= 0, so f * 2 + lo >= lo already holds after rounding and the clamp is the identity.
    // (One FMA on purpose: the producers of the warp-specialised kernel run one dependent chain per warp, and
    // splitting this or the first Horner step of the log into mul + add to save the literal moves cost 4 %.)
    return fma(f, 2.0, SDE_LOGPOLY[5]);
}

__device__ __forceinline__ float uniform_pm1(u32 bits) {
    const float lo = -0.99999994f;  // nextafter(-1f, 0f)
    float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
    return fmaf(f, 2.0f, lo);   // the clamp max(lo, .) is the identity here, see the double version
}

// N standard normals at once: out[i] = element e[i] of jax.random.normal(key[i], shape with size[i]
// elements, dtype).  Every stage is written group-wide so that the N dependent chains (20 threefry
// rounds, the log, the 22-step Horner recurrence) are interleaved in the instruction stream.
// The two halves of a draw, separately: the uniform variate in (-1, 1) (threefry2x32 + mantissa fill: integer work) and
// sqrt(2) erf_inv of it (floating-point work).  normal_group below is their composition; the warp-specialised kernel can
// also run them in different warps (SDE_WS_SPLIT_NORMAL: producers draw uniforms, consumers turn them into normals).
__device__ __forceinline__ real uniform_elem(const Key key, const u32 e, const u32 size) {
#if SDE_F64
    return uniform_pm1(random_bits64(key, e, size));
#else
    return uniform_pm1(random_bits32(key, e, size));
#endif
}

template <int N>
__device__ __forceinline__ void normals_from_uniforms(const real (&x)[N], real (&out)[N]) {
    real w[N], p[N];
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = erfinv_w(x[i]);
#if SDE_F64
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = ERFINV64_LT6_25[0];
#pragma unroll
    for (int c = 1; c < 23; ++c) {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = fma(p[i], w[i] - 3.125, ERFINV64_LT6_25[c]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        real v = p[i] * x[i];
        if (!erfinv_is_central(w[i])) v = erfinv_tail(x[i], w[i]);
        out[i] = SDE_LOGPOLY[3] * v;
    }
#else
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = ERFINV32_LT5[0];
#pragma unroll
    for (int c = 1; c < 9; ++c) {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = fmaf(p[i], w[i] - 2.5f, ERFINV32_LT5[c]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        real v = p[i] * x[i];
        if (!erfinv_is_central(w[i])) v = erfinv_tail(x[i], w[i]);
        out[i] = 1.41421354f * v;
    }
#endif
}

template <int N>
__device__ __forceinline__ void normal_group(const Key (&key)[N], const u32 (&e)[N], const u32 (&size)[N],
                                             real (&out)[N]) {
    real x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = uniform_elem(key[i], e[i], size[i]);
    normals_from_uniforms<N>(x, out);
}

// one standard normal: element e of jax.random.normal(key, shape with `size` elements, dtype)
__device__ __forceinline__ real normal_elem(Key key, u32 e, u32 size) {
    const Key k[1] = {key};
    const u32 ee[1] = {e}, ss[1] = {size};
    real o[1];
    normal_group<1>(k, ee, ss, o);
    return o[0];
}

// ----------------------------------------------------------------------------------------
// Integrals of one step, already scaled by the powers of dt (reference sde_solver.py:233-241)
// ----------------------------------------------------------------------------------------
struct Integrals {
    real dW[SDE_M];
#if SDE_SCHEME >= SDE_MILSTEIN
    real Iww[SDE_M][SDE_M];             // I_(j1,j2)
#endif
#if SDE_SCHEME >= SDE_WAGNER_PLATEN
    real Iwt[SDE_M];                    // I_(j,0)
    real Itw[SDE_M];                    // I_(0,j)
    real Iwww[SDE_M][SDE_M][SDE_M];     // I_(j1,j2,j3)
#endif
};

struct StepScale {
    real sqrt_dt;   // sqrt(dt) in the working dtype
    real dt;
    real dt32;      // dt**1.5 computed in double on the host, cast
};

// Number of normals per step that are produced by a ROLLED loop and staged through this thread's
// column of shared memory before being read back with compile-time indices.  Keeping the
// normal generator (threefry + log1p + 23-term Horner) out of the unrolled code is what keeps the
// step body inside the instruction cache.
#if SDE_M > 1 && SDE_SCHEME == SDE_WAGNER_PLATEN
#define SDE_STAGE_COUNT (2 * SDE_M * SDE_P + 3 * SDE_M)
#elif SDE_M > 4 && SDE_SCHEME == SDE_EULER
#define SDE_STAGE_COUNT (SDE_M)
#else
#define SDE_STAGE_COUNT 0
#endif
#ifndef SDE_BLOCK
#define SDE_BLOCK 128
#endif
// Steps whose integrals are generated together (see WienerGen::generate_block).
#if SDE_M == 1
#define SDE_STEP_BLOCK ((SDE_SCHEME == SDE_WAGNER_PLATEN) ? 2 : 4)
#elif SDE_SCHEME == SDE_EULER && SDE_STAGE_COUNT == 0
#define SDE_STEP_BLOCK ((SDE_M >= 4) ? 1 : ((SDE_M == 3) ? 1 : 2))
#else
#define SDE_STEP_BLOCK 1
#endif
// element k of this thread's staging column (stage points at the thread's slot of row 0)
#ifndef SDE_STAGE_STRIDE
#define SDE_STAGE_STRIDE SDE_BLOCK   // threads sharing one staging row
#endif
#define SDE_STAGE(k) stage[(k) * SDE_STAGE_STRIDE]

// compile-time helpers ---------------------------------------------------------------------
#if SDE_P <= 16
#define SDE_UNROLL _Pragma("unroll")
#else
#define SDE_UNROLL _Pragma("unroll 1")
#endif

// out[l] += sgn * sum_{r=1}^{P-l} X[r] * Y[l+r]      (l = 1..P-1)   arrays indexed 1..P
__device__ __forceinline__ void acc_corr(const real* X, const real* Y, real* out, const real sgn) {
    SDE_UNROLL
    for (int l = 1; l < SDE_P; ++l) {
        real s = out[l];
        SDE_UNROLL
        for (int r = 1; r <= SDE_P - l; ++r) s = fma(sgn * X[r], Y[l + r], s);
        out[l] = s;
    }
}

// out[l] += sgn * sum_{r=1}^{l-1} X[r] * Y[l-r]      (l = 2..P)
__device__ __forceinline__ void acc_conv(const real* X, const real* Y, real* out, const real sgn) {
    SDE_UNROLL
    for (int l = 2; l <= SDE_P; ++l) {
        real s = out[l];
        SDE_UNROLL
        for (int r = 1; r < l; ++r) s = fma(sgn * X[r], Y[l - r], s);
        out[l] = s;
    }
}

__device__ __forceinline__ void zero_arr(real* a) {
    SDE_UNROLL
    for (int l = 0; l <= SDE_P; ++l) a[l] = RC(0.0);
}

// ----------------------------------------------------------------------------------------
// WienerGen: per-fragment generator.  init(fragment_key, S) once per fragment, then
// generate(s, scale, I) for step s in [0, S) of that fragment.
// ----------------------------------------------------------------------------------------
struct WienerGen {
    u32 S;
#if SDE_M == 1 || SDE_SCHEME == SDE_EULER
    Key k0;
#elif SDE_SCHEME == SDE_MILSTEIN
    Key k_xi, k_mu, k_eta, k_zeta;
#else
    Key k_xi, k_zeta, k_eta, k_mu, k_phi;
#endif

    __device__ __forceinline__ void init(Key fkey, u32 steps) {
        S = steps;
#if SDE_M == 1 || SDE_SCHEME == SDE_EULER
        k0 = fkey;
#elif SDE_SCHEME == SDE_MILSTEIN
        k_xi = split_key(fkey, 4, 0);
        k_mu = split_key(fkey, 4, 1);
        k_eta = split_key(fkey, 4, 2);
        k_zeta = split_key(fkey, 4, 3);
#else
        k_xi = split_key(fkey, 5, 0);
        k_zeta = split_key(fkey, 5, 1);
        k_eta = split_key(fkey, 5, 2);
        k_mu = split_key(fkey, 5, 3);
        k_phi = split_key(fkey, 5, 4);
#endif
    }

    // Integrals of SDE_STEP_BLOCK consecutive steps s0, s0+1, ... (steps past the end of the fragment produce
    // unused values).  Blocking several steps of a cheap scheme into one group of normals gives the dependent
    // threefry / Horner chains something to interleave with.
    __device__ __forceinline__ void generate_block(u32 s0, const StepScale& sc, Integrals (&I)[SDE_STEP_BLOCK],
                                                   real* stage) const {
#if SDE_M == 1
        // closed forms, reference vectorized_I_generation.py:265-294
#if SDE_SCHEME == SDE_WAGNER_PLATEN
        constexpr int B = SDE_STEP_BLOCK;
        Key kk[2 * B];
        u32 ee[2 * B], ss[2 * B];
        real nn[2 * B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            kk[b] = k0; kk[B + b] = k0;
            ee[b] = s0 + b; ee[B + b] = S + s0 + b;
            ss[b] = ss[B + b] = 2u * S;
        }
        normal_group<2 * B>(kk, ee, ss, nn);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const real u0 = nn[b], u1 = nn[B + b];
            const real dZ = RC(0.5) * (u0 + RC(0.57735026918962576451) * u1);
            I[b].dW[0] = sc.sqrt_dt * u0;
            I[b].Iww[0][0] = sc.dt * (RC(0.5) * (u0 * u0 - RC(1.0)));
            I[b].Iwt[0] = sc.dt32 * dZ;
            I[b].Itw[0] = sc.dt32 * (u0 - dZ);
            I[b].Iwww[0][0][0] = sc.dt32 * (RC(0.5) * (RC(0.33333333333333333333) * u0 * u0 - RC(1.0)) * u0);
        }
#else
        constexpr int B = SDE_STEP_BLOCK;
        Key kk[B];
        u32 ee[B], ss[B];
        real nn[B];
#pragma unroll
        for (int b = 0; b < B; ++b) { kk[b] = k0; ee[b] = s0 + b; ss[b] = S; }
        normal_group<B>(kk, ee, ss, nn);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            I[b].dW[0] = sc.sqrt_dt * nn[b];
#if SDE_SCHEME == SDE_MILSTEIN
            I[b].Iww[0][0] = sc.dt * (RC(0.5) * (nn[b] * nn[b] - RC(1.0)));
#endif
        }
#endif
#elif SDE_SCHEME == SDE_EULER
#if SDE_STAGE_COUNT > 0
        {
            // rolled loop over groups of G independent draws (G = 4, 3 or 2 when it divides M): the dependent
            // threefry / Horner chains of a group interleave, the loop body stays small
#ifndef SDE_EULER_GROUP
#define SDE_EULER_GROUP ((SDE_M % 4 == 0) ? 4 : ((SDE_M % 3 == 0) ? 3 : ((SDE_M % 2 == 0) ? 2 : 1)))
#endif
            constexpr int G = SDE_EULER_GROUP;
            static_assert(SDE_M % G == 0, "SDE_EULER_GROUP must divide SDE_M");
#pragma unroll 1
            for (int j = 0; j < SDE_M; j += G) {
                Key kk[G];
                u32 ee[G], ss[G];
                real nn[G];
#pragma unroll
                for (int q = 0; q < G; ++q) { kk[q] = k0; ee[q] = s0 * SDE_M + j + q; ss[q] = S * SDE_M; }
                normal_group<G>(kk, ee, ss, nn);
#pragma unroll
                for (int q = 0; q < G; ++q) SDE_STAGE(j + q) = nn[q];
            }
        }
#pragma unroll
        for (int j = 0; j < SDE_M; ++j) I[0].dW[j] = sc.sqrt_dt * SDE_STAGE(j);
#else
        {
            constexpr int N = SDE_M * SDE_STEP_BLOCK;     // elements of consecutive steps are consecutive
            Key kk[N];
            u32 ee[N], ss[N];
            real nn[N];
#pragma unroll
            for (int q = 0; q < N; ++q) { kk[q] = k0; ee[q] = s0 * SDE_M + q; ss[q] = S * SDE_M; }
            normal_group<N>(kk, ee, ss, nn);
#pragma unroll
            for (int b = 0; b < SDE_STEP_BLOCK; ++b)
#pragma unroll
                for (int j = 0; j < SDE_M; ++j) I[b].dW[j] = sc.sqrt_dt * nn[b * SDE_M + j];
        }
#endif
#elif SDE_SCHEME == SDE_MILSTEIN
        gen_milstein(s0, sc, I[0]);
#else
        gen_wagner_platen(s0, sc, I[0], stage);
#endif
        (void)stage;
    }

#if SDE_M > 1 && SDE_SCHEME == SDE_MILSTEIN
    // reference vectorized_I_generation.py:301-339 (Kloeden-Platen 10.3.7, truncated at p)
    __device__ __forceinline__ void gen_milstein(u32 s, const StepScale& sc, Integrals& I) const {
        constexpr int M = SDE_M, P = SDE_P;
        real xi[M], mu[M];
        {
            Key kk[2 * M];
            u32 ee[2 * M], ss[2 * M];
            real nn[2 * M];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                kk[j] = k_xi; kk[M + j] = k_mu;
                ee[j] = ee[M + j] = s * M + j;
                ss[j] = ss[M + j] = S * M;
            }
            normal_group<2 * M>(kk, ee, ss, nn);
#pragma unroll
            for (int j = 0; j < M; ++j) { xi[j] = nn[j]; mu[j] = nn[M + j]; }
        }
        real ser[M][M];  // upper triangle: sum_r (1/r) (zeta_ri a_rj - a_ri zeta_rj)
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) ser[i][j] = RC(0.0);
#pragma unroll 1
        for (int r = 1; r <= P; ++r) {
            real a[M], z[M];
            {
                Key kk[2 * M];
                u32 ee[2 * M], ss[2 * M];
                real nn[2 * M];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    kk[j] = k_eta; kk[M + j] = k_zeta;
                    ee[j] = ee[M + j] = (s * P + (r - 1)) * M + j;
                    ss[j] = ss[M + j] = S * P * M;
                }
                normal_group<2 * M>(kk, ee, ss, nn);
#pragma unroll
                for (int j = 0; j < M; ++j) { a[j] = fma(RC(1.4142135623730950488), xi[j], nn[j]); z[j] = nn[M + j]; }
            }
            const real rec = RC(1.0) / (real)r;
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = i + 1; j < M; ++j) ser[i][j] = fma(rec, z[i] * a[j] - a[i] * z[j], ser[i][j]);
        }
        real rho = RC(0.0);
        SDE_UNROLL
        for (int r = 1; r <= P; ++r) rho += RC(1.0) / (real)(r * r);
        rho = RC(0.083333333333333333333) - rho * (real)(0.5 / (SDE_PI * SDE_PI));  // 1/12 - sum/(2 pi^2)
        const real sqrt_rho = sqrt(rho);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            I.dW[i] = sc.sqrt_dt * xi[i];
            I.Iww[i][i] = sc.dt * (RC(0.5) * (xi[i] * xi[i] - RC(1.0)));
#pragma unroll
            for (int j = i + 1; j < M; ++j) {
                const real sym = RC(0.5) * xi[i] * xi[j];
                const real asym = sqrt_rho * (xi[i] * mu[j] - mu[i] * xi[j]) +
                                  (real)(0.5 / SDE_PI) * ser[i][j];
                I.Iww[i][j] = sc.dt * (sym + asym);
                I.Iww[j][i] = sc.dt * (sym - asym);
            }
        }
    }
#endif

#if SDE_M > 1 && SDE_SCHEME == SDE_WAGNER_PLATEN
    // reference vectorized_I_generation.py:340-470 (Kloeden-Platen 5.8.11 / 10.4.7)
    // Works on the prescaled series coefficients zs[j][r] = zeta_{r,j}/r, es[j][r] = eta_{r,j}/r,
    // which turns every term of the C and D double sums into a single FMA, and shares the
    // correlation / convolution partial sums between D[j1,.,j3] and D[j3,.,j1].
    __device__ __forceinline__ void gen_wagner_platen(u32 s, const StepScale& sc, Integrals& I, real* stage) const {
        wp_draw(s, stage);
        wp_assemble(sc, I, stage, [] {});
    }

    // (key, element, size) of staged normal n of step s: n < MP zeta, < 2MP eta, then xi, mu, phi (M each)
    __device__ __forceinline__ void wp_item(u32 s, int n, Key& key, u32& e, u32& size) const {
        constexpr int M = SDE_M, P = SDE_P;
        if (n < 2 * P * M) {
            const bool first = n < P * M;
            key = first ? k_zeta : k_eta;
            e = s * (u32)(P * M) + (u32)(first ? n : n - P * M);
            size = S * (u32)(P * M);
        } else {
            const int q = n - 2 * P * M;
            key = (q < M) ? k_xi : ((q < 2 * M) ? k_mu : k_phi);
            e = s * M + (u32)(q % M);
            size = S * M;
        }
    }

    // -- normals of step s: rolled loops into the staging column
    __device__ __forceinline__ void wp_draw(u32 s, real* stage) const {
        constexpr int M = SDE_M, P = SDE_P;
        // 2M independent normals per iteration: enough instruction-level parallelism to cover the
        // dependent threefry rounds / Horner chains with only two resident warps per scheduler.
        {
            const u32 base = s * (u32)(P * M), size = S * (u32)(P * M);
#ifndef SDE_NORMAL_ROWS
#define SDE_NORMAL_ROWS 1          // series terms r generated together (2M normals each)
#endif
            static_assert(P % SDE_NORMAL_ROWS == 0, "SDE_P must be a multiple of SDE_NORMAL_ROWS");
#pragma unroll 1
            for (int r = 0; r < P; r += SDE_NORMAL_ROWS) {
                constexpr int G = SDE_NORMAL_ROWS * M;
                Key kk[2 * G];
                u32 ee[2 * G], ss[2 * G];
                real nn[2 * G];
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    kk[q] = k_zeta; kk[G + q] = k_eta;
                    ee[q] = ee[G + q] = base + r * M + q;
                    ss[q] = ss[G + q] = size;
                }
                normal_group<2 * G>(kk, ee, ss, nn);
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    SDE_STAGE(r * M + q) = nn[q];
                    SDE_STAGE(P * M + r * M + q) = nn[G + q];
                }
            }
#pragma unroll 1
            for (int j = 0; j < M; ++j) {
                const Key kk[3] = {k_xi, k_mu, k_phi};
                const u32 ee[3] = {s * M + j, s * M + j, s * M + j}, ss[3] = {S * M, S * M, S * M};
                real nn[3];
                normal_group<3>(kk, ee, ss, nn);
                SDE_STAGE(2 * P * M + j) = nn[0];
                SDE_STAGE(2 * P * M + M + j) = nn[1];
                SDE_STAGE(2 * P * M + 2 * M + j) = nn[2];
            }
        }
    }

    // -- staged normals -> iterated integrals of one step; `after_reads()` is invoked once every staged value
    //    has been copied into registers (the warp-specialised kernel releases the ring stage there)
    //    The last CT of the 3M values (xi, mu, phi) may come from the caller's registers (`own`) instead of the stage.
    template <int CT = 0, class AfterReads>
    __device__ __forceinline__ void wp_assemble(const StepScale& sc, Integrals& I, const real* stage,
                                                AfterReads after_reads, const real* own = nullptr) const {
        constexpr int M = SDE_M, P = SDE_P;
#define SDE_TAIL(q) (((q) >= 3 * M - CT) ? own[(q) - (3 * M - CT)] : SDE_STAGE(2 * P * M + (q)))
        real xi[M], a[M], b[M];
        real zs[M][P + 1], es[M][P + 1];
        real sa[M];                       // sum_r zs[j][r], reused by the C matrix
        {
            real sb[M];
#pragma unroll
            for (int j = 0; j < M; ++j) { sa[j] = RC(0.0); sb[j] = RC(0.0); zs[j][0] = RC(0.0); es[j][0] = RC(0.0); }
            SDE_UNROLL
            for (int r = 1; r <= P; ++r) {
                const real rec = RC(1.0) / (real)r;
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    const real z = rec * SDE_STAGE((r - 1) * M + j);
                    const real h = rec * SDE_STAGE(P * M + (r - 1) * M + j);
                    zs[j][r] = z;
                    es[j][r] = h;
                    sa[j] += z;                       // sum zeta_r / r
                    sb[j] = fma(rec, h, sb[j]);       // sum eta_r / r^2
                }
            }
            real rho = RC(0.0), alpha = RC(0.0);
            SDE_UNROLL
            for (int r = 1; r <= P; ++r) {
                const real r2 = RC(1.0) / (real)(r * r);
                rho += r2;
                alpha += r2 * r2;
            }
            rho = RC(0.083333333333333333333) - rho * (real)(0.5 / (SDE_PI * SDE_PI));      // 1/12 - sum/(2 pi^2)
            alpha = (real)(SDE_PI * SDE_PI / 180.0) - alpha * (real)(0.5 / (SDE_PI * SDE_PI));  // pi^2/180 - sum/(2 pi^2)
            const real two_sqrt_rho = RC(2.0) * sqrt(rho), sqrt_alpha = sqrt(alpha);
#pragma unroll
            for (int j = 0; j < M; ++j) {
                xi[j] = SDE_TAIL(j);
                const real mu = SDE_TAIL(M + j);
                const real phi = SDE_TAIL(2 * M + j);
#undef SDE_TAIL
                a[j] = -(real)(1.4142135623730950488 / SDE_PI) * sa[j] - two_sqrt_rho * mu;   // sqrt(2)/pi
                b[j] = fma((real)(1.0 / (1.4142135623730950488 * SDE_PI)), sb[j], sqrt_alpha * phi);  // 1/(sqrt(2) pi)
            }
            after_reads();
        }

        // A (antisymmetric), B (symmetric); the zeta-zeta dot products are kept for C
        real A[M][M], B[M][M], dzz[M][M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
#pragma unroll
            for (int j = i; j < M; ++j) {
                real sA = RC(0.0), sZ = RC(0.0), sE = RC(0.0);
                SDE_UNROLL
                for (int r = 1; r <= P; ++r) {
                    sZ = fma(zs[i][r], zs[j][r], sZ);
                    sE = fma(es[i][r], es[j][r], sE);
                    if (j > i) sA = fma((real)r * zs[i][r], es[j][r], fma(-(real)r * es[i][r], zs[j][r], sA));
                }
                dzz[i][j] = dzz[j][i] = sZ;
                B[i][j] = B[j][i] = (real)(0.25 / (SDE_PI * SDE_PI)) * (sZ + sE);     // 1/(4 pi^2)
                A[i][j] = (real)(0.5 / SDE_PI) * sA;                // 1/(2 pi)
                A[j][i] = -A[i][j];
            }
        }

        // C[j1][j2] = -1/(2 pi^2) sum_{l != r} [ r^2/(r^2-l^2) zs[j1][r] zs[j2][l] + r l/(r^2-l^2) es[j1][r] es[j2][l] ]
        // Both kernels split into a constant and an antisymmetric part: r^2/(r^2-l^2) = 1/2 + (r^2+l^2)/(2(r^2-l^2)),
        // r l/(r^2-l^2) = -(l r/(l^2-r^2)).  The constant part is 1/2 (S_j1 S_j2 - zs_j1 . zs_j2) with S_j = sum_r zs[j][r];
        // the antisymmetric parts are sums over l < r of 2x2 determinants, change sign under j1 <-> j2 and vanish
        // for j1 = j2 - so C costs one pass over the 45 pairs per unordered (j1, j2) instead of 2 x 90 terms per j1.
        real C[M][M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            C[i][i] = -(real)(0.25 / (SDE_PI * SDE_PI)) * (sa[i] * sa[i] - dzz[i][i]);
#pragma unroll
            for (int j = i + 1; j < M; ++j) {
                real asym = RC(0.0);
                SDE_UNROLL
                for (int l = 1; l < P; ++l) {
                    SDE_UNROLL
                    for (int r = l + 1; r <= P; ++r) {
                        const real dz = fma(-zs[i][l], zs[j][r], zs[i][r] * zs[j][l]);
                        const real de = fma(-es[i][l], es[j][r], es[i][r] * es[j][l]);
                        asym = fma((real)(r * r + l * l) / (real)(2 * (r * r - l * l)), dz, asym);
                        asym = fma((real)(r * l) / (real)(r * r - l * l), de, asym);
                    }
                }
                const real sym = RC(0.5) * (sa[i] * sa[j] - dzz[i][j]);
                C[i][j] = -(real)(0.5 / (SDE_PI * SDE_PI)) * (sym + asym);
                C[j][i] = -(real)(0.5 / (SDE_PI * SDE_PI)) * (sym - asym);
            }
        }

        // D[j1][j2][j3] = K sum_l ( zs[j2][l] Tz_{j1 j3}[l] + es[j2][l] Te_{j1 j3}[l] ), the factor l
        // that undoes the prescaling of the j2 coefficient is folded into Tz / Te.
        real D[M][M][M];
        const real KD = (real)(1.0 / (SDE_PI * SDE_PI * 5.6568542494923801952));  // 1 / (pi^2 2^(5/2))
#pragma unroll
        for (int ja = 0; ja < M; ++ja) {
            {   // j1 = j3 = ja:  Tz = 2 (corr_ze - corr_ez + conv_ze),  Te = conv_ee - conv_zz
                real Tz[P + 1], Te[P + 1];
                zero_arr(Tz);
                zero_arr(Te);
                acc_corr(zs[ja], es[ja], Tz, RC(1.0));
                acc_corr(es[ja], zs[ja], Tz, -RC(1.0));
                acc_conv(zs[ja], es[ja], Tz, RC(1.0));
                acc_conv(es[ja], es[ja], Te, RC(1.0));
                acc_conv(zs[ja], zs[ja], Te, -RC(1.0));
                SDE_UNROLL
                for (int l = 1; l <= P; ++l) {
                    Tz[l] *= (real)(2 * l);
                    Te[l] *= (real)l;
                }
#pragma unroll
                for (int j2 = 0; j2 < M; ++j2) {
                    real d = RC(0.0);
                    SDE_UNROLL
                    for (int l = 1; l <= P; ++l) d = fma(zs[j2][l], Tz[l], fma(es[j2][l], Te[l], d));
                    D[ja][j2][ja] = KD * d;
                }
            }
#pragma unroll
            for (int jb = ja + 1; jb < M; ++jb) {
                real dab[M], dba[M];
                {
                    // Tz(a,b) = Q + 2 corr_ze[b,a],  Tz(b,a) = Q + 2 corr_ze[a,b]
                    // Q = conv_ze[a,b] + conv_ze[b,a] - corr_ez[a,b] - corr_ez[b,a]
                    real Q[P + 1], R1[P + 1], R2[P + 1];
                    zero_arr(Q);
                    zero_arr(R1);
                    zero_arr(R2);
                    acc_corr(es[ja], zs[jb], Q, -RC(1.0));
                    acc_corr(es[jb], zs[ja], Q, -RC(1.0));
                    acc_conv(zs[ja], es[jb], Q, RC(1.0));
                    acc_conv(zs[jb], es[ja], Q, RC(1.0));
                    acc_corr(zs[jb], es[ja], R1, RC(1.0));
                    acc_corr(zs[ja], es[jb], R2, RC(1.0));
                    SDE_UNROLL
                    for (int l = 1; l <= P; ++l) {
                        const real ql = (real)l * Q[l];
                        R1[l] = fma((real)(2 * l), R1[l], ql);
                        R2[l] = fma((real)(2 * l), R2[l], ql);
                    }
#pragma unroll
                    for (int j2 = 0; j2 < M; ++j2) {
                        real d1 = RC(0.0), d2 = RC(0.0);
                        SDE_UNROLL
                        for (int l = 1; l <= P; ++l) {
                            d1 = fma(zs[j2][l], R1[l], d1);
                            d2 = fma(zs[j2][l], R2[l], d2);
                        }
                        dab[j2] = d1;
                        dba[j2] = d2;
                    }
                }
                {
                    // Te(a,b) = W + V, Te(b,a) = W - V
                    // V = corr_zz[b,a] - corr_zz[a,b] + corr_ee[b,a] - corr_ee[a,b],  W = conv_ee[a,b] - conv_zz[a,b]
                    real V[P + 1], W[P + 1];
                    zero_arr(V);
                    zero_arr(W);
                    acc_corr(zs[jb], zs[ja], V, RC(1.0));
                    acc_corr(zs[ja], zs[jb], V, -RC(1.0));
                    acc_corr(es[jb], es[ja], V, RC(1.0));
                    acc_corr(es[ja], es[jb], V, -RC(1.0));
                    acc_conv(es[ja], es[jb], W, RC(1.0));
                    acc_conv(zs[ja], zs[jb], W, -RC(1.0));
                    SDE_UNROLL
                    for (int l = 1; l <= P; ++l) {
                        const real pw = (real)l * W[l], qv = (real)l * V[l];
                        V[l] = pw + qv;
                        W[l] = pw - qv;
                    }
#pragma unroll
                    for (int j2 = 0; j2 < M; ++j2) {
                        real d1 = dab[j2], d2 = dba[j2];
                        SDE_UNROLL
                        for (int l = 1; l <= P; ++l) {
                            d1 = fma(es[j2][l], V[l], d1);
                            d2 = fma(es[j2][l], W[l], d2);
                        }
                        D[ja][j2][jb] = KD * d1;
                        D[jb][j2][ja] = KD * d2;
                    }
                }
            }
        }

        // assembly, reference vectorized_I_generation.py:393-470
        real wt[M], tw[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
            wt[j] = RC(0.5) * (xi[j] + a[j]);
            tw[j] = xi[j] - wt[j];
            I.dW[j] = sc.sqrt_dt * xi[j];
            I.Iwt[j] = sc.dt32 * wt[j];
            I.Itw[j] = sc.dt32 * tw[j];
        }
        real Jww[M][M], Jtww[M][M];
        const real inv_pi = (real)(1.0 / SDE_PI);
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const real xx = xi[i] * xi[j];
                Jww[i][j] = RC(0.5) * xx - RC(0.5) * (xi[i] * a[j] - xi[j] * a[i]) + A[i][j];
                Jtww[i][j] = RC(0.16666666666666666667) * xx - inv_pi * xi[j] * b[i] + B[i][j] -
                             RC(0.25) * a[j] * xi[i] + RC(0.5) * inv_pi * xi[i] * b[j] + C[i][j] +
                             RC(0.5) * A[i][j];
                I.Iww[i][j] = sc.dt * ((i == j) ? RC(0.5) * (xx - RC(1.0)) : Jww[i][j]);
            }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j)
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    real J = xi[i] * Jtww[j][k] + RC(0.5) * a[i] * Jww[j][k] +
                             RC(0.5) * inv_pi * b[i] * xi[j] * xi[k] - xi[j] * B[i][k] +
                             xi[k] * (RC(0.5) * A[i][j] - C[j][i]) + D[i][j][k];
                    if (i == j) J -= RC(0.5) * tw[k];
                    if (j == k) J -= RC(0.5) * wt[i];
                    I.Iwww[i][j][k] = sc.dt32 * J;
                }
    }
#endif
};

}  // namespace sdeb
