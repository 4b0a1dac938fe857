// Per-row softmax arithmetic shared by the TMEM-ring attention kernels (attention7.cuh, attention8.cuh): 32-key chunk
// helpers (row maximum, masking of non-existent keys, exponentiation + row sum + fp16 packing) and the rare-path rescale
// of O / P in tensor memory.  Semantics: reference soft_max_ext over scaled scores (dinov2.cpp:527-543, ops.cpp:4641-4737),
// evaluated as exp2((s - m) * log2(e)/8) with a lazily updated reference maximum m.
#pragma once
#include "ptx.cuh"

// every ATS_POLY_MOD-th pair of probabilities is computed with a polynomial on the FMA pipe instead of MUFU (0 = none)
// Round 2: with one MMA-issuing warp per query tile (attention10.cuh) the exponentiation phase runs the XU pipe (MUFU.EX2 at
// 16 / clk / SM plus the F2FP packs) at ~94 %, so moving exponentials to the FMA pipe pays again.  Per ViT-L layer (B = 64) on
// v10: 0 -> 705 us, 8 -> 697, 4 -> 675, 3 -> 666, 2 -> 694 (a polynomial pair costs ~14 issue slots against 6 for a MUFU pair).
#ifndef ATS_POLY_MOD
#define ATS_POLY_MOD 3
#endif

namespace dino {

// exp2(x) without the MUFU unit: round-to-nearest split x = n + f (magic-number add), cubic minimax for 2^f on
// [-0.5, 0.5], exponent field patched by integer add.  Arguments below -30 (masked keys are -inf) clamp to 2^-30, which
// is zero once P is rounded to fp16.
__device__ __forceinline__ float attn_exp2_poly3(float x) {
    const float t = fmaxf(x, -30.0f);
    const float u = t + 12582912.0f;                 // 1.5 * 2^23: low mantissa bits now hold round(t)
    const float f = t - (u - 12582912.0f);
    float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
    p = fmaf(p, f, 0.6932762265205383f);
    p = fmaf(p, f, 0.9999289512634277f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(u) << 23));
}

// Rare path of the lazy running-max correction: scale this thread's row of O_t (64 fp32 columns) in TMEM and, when the
// growth was found in the middle of a tile, the first p_cols packed columns of P it has already written.  Inlined and
// rolled (8 columns at a time): a call here would make ptxas save the 64-128 live score registers on EVERY tile.
__device__ __forceinline__ void attn_rescale(uint32_t o_addr, uint32_t p_addr, float alpha, bool do_o, int p_cols) {
    tmem_st_wait();                                       // this thread's earlier P stores have landed
    if (do_o) {
#pragma unroll 1
        for (int i = 0; i < 64; i += 8) {
            uint32_t a[8];
            tmem_ld_32x32b_x8(o_addr + i, a);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 8; ++d) a[d] = __float_as_uint(__uint_as_float(a[d]) * alpha);
            tmem_st_32x32b_x8(o_addr + i, a);
        }
    }
    if (p_cols > 0) {
        const __half2 a2 = __float2half2_rn(alpha);
#pragma unroll 1
        for (int i = 0; i < p_cols; i += 8) {
            uint32_t q[8];
            tmem_ld_32x32b_x8(p_addr + i, q);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                __half2 v = *reinterpret_cast<__half2 *>(&q[d]);
                v = __hmul2(v, a2);
                q[d] = *reinterpret_cast<uint32_t *>(&v);
            }
            tmem_st_32x32b_x8(p_addr + i, q);
        }
    }
    tmem_st_wait();
}

__device__ __forceinline__ float attn_rowmax32(const uint32_t (&v)[32]) {
    float m0 = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    float m1 = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
#pragma unroll
    for (int i = 6; i < 30; i += 4) {
        m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    }
    return fmax3(fmax3(m0, __uint_as_float(v[30]), __uint_as_float(v[31])), m1, m1);
}
// keys at or beyond `valid` (relative to the chunk) do not exist: -inf
__device__ __forceinline__ void attn_mask32(uint32_t (&v)[32], int valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i >= valid) v[i] = 0xFF800000u;
}
// pairs [E0, E1) of a 32-key chunk: p = exp2(s c - mc) (MUFU, every ATS_POLY_MOD-th pair on the FMA pipe), fp32 row sum,
// one rounding to packed fp16.  ATS_PACKED: the scale-and-shift, the row sum and the polynomial run on packed fp32 pairs
// (FFMA2 / FADD2: one issue slot per two keys).
#ifndef ATS_PACKED
#define ATS_PACKED 1   // measured on v8, per ViT-L layer (B=64): scalar 769 us; packed 724 us; packed + 1/4 polynomial 725, 1/3 761, 1/2 768
#endif
// ATS_PIPE: the row-sum add and the fp16 pack of pair e are placed ATS_PIPE pairs after its two MUFU ops in program order.
// Measured (SASS of ATS_PIPE = 1, 2, 3 is identical): ptxas re-schedules the block on its own and always pipelines by exactly
// one pair — MUFU (stall 8), MUFU, FADD2(previous pair), F2FP(previous pair, stall 6) — whatever the source order says.
#ifndef ATS_PIPE
#define ATS_PIPE 1
#endif
// 1: the MUFU and row-sum instructions are volatile asm (fixed order among themselves at the NVVM level)
#ifndef ATS_VOLATILE
#define ATS_VOLATILE 1
#endif
template <int E0, int E1>
__device__ __forceinline__ void attn_exp_pairs(const uint32_t (&v)[32], uint32_t (&pk)[16], float c, float mc, float (&ls)[2]) {
#if ATS_PACKED
    const f32x2 c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
    f32x2 acc = pack_f32x2(ls[0], ls[1]);
    float q0[E1 - E0], q1[E1 - E0];
#pragma unroll
    for (int e = E0; e < E1 + ATS_PIPE; ++e) {
        if (e < E1) {
            const f32x2 x = fma2_f32(pack_f32x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), c2, nmc2);
            float p0, p1;
            if (ATS_POLY_MOD > 0 && (e % (ATS_POLY_MOD > 0 ? ATS_POLY_MOD : 1)) == ATS_POLY_MOD - 1) {
                // exp2 on the FMA pipe, two keys per instruction: clamp, round-to-nearest split x = n + f (magic-number add), cubic
                // for 2^f on [-0.5, 0.5], exponent patched in with one LEA per key
                float x0, x1;
                unpack_f32x2(x, x0, x1);
                const f32x2 t = pack_f32x2(fmaxf(x0, -30.0f), fmaxf(x1, -30.0f));
                const f32x2 magic = pack_f32x2(12582912.0f, 12582912.0f), nmagic = pack_f32x2(-12582912.0f, -12582912.0f);
                const f32x2 u = add2_f32(t, magic);
                const f32x2 w = add2_f32(u, nmagic);
                float w0, w1;
                unpack_f32x2(w, w0, w1);
                const f32x2 f = add2_f32(t, pack_f32x2(-w0, -w1));
                f32x2 q = fma2_f32(pack_f32x2(0.05508868396282196f, 0.05508868396282196f), f, pack_f32x2(0.24260404706001282f, 0.24260404706001282f));
                q = fma2_f32(q, f, pack_f32x2(0.6932762265205383f, 0.6932762265205383f));
                q = fma2_f32(q, f, pack_f32x2(0.9999289512634277f, 0.9999289512634277f));
                float r0, r1, u0, u1;
                unpack_f32x2(q, r0, r1);
                unpack_f32x2(u, u0, u1);
                p0 = __int_as_float(__float_as_int(r0) + (__float_as_int(u0) << 23));
                p1 = __int_as_float(__float_as_int(r1) + (__float_as_int(u1) << 23));
            } else {
                float x0, x1;
                unpack_f32x2(x, x0, x1);
#if ATS_VOLATILE
                p0 = ex2_approx_ordered(x0);
                p1 = ex2_approx_ordered(x1);
#else
                p0 = ex2_approx(x0);
                p1 = ex2_approx(x1);
#endif
            }
            q0[e - E0] = p0;
            q1[e - E0] = p1;
        }
        if (e - ATS_PIPE >= E0) {
            const int d = e - ATS_PIPE;
#if ATS_VOLATILE
            acc = add2_f32_ordered(acc, pack_f32x2(q0[d - E0], q1[d - E0]));
#else
            acc = add2_f32(acc, pack_f32x2(q0[d - E0], q1[d - E0]));
#endif
            pk[d] = cvt_f16x2(q0[d - E0], q1[d - E0]);
        }
    }
    unpack_f32x2(acc, ls[0], ls[1]);
#else
#pragma unroll
    for (int e = E0; e < E1; ++e) {
        const float x0 = fmaf(__uint_as_float(v[2 * e]), c, -mc);
        const float x1 = fmaf(__uint_as_float(v[2 * e + 1]), c, -mc);
        float p0, p1;
        if (ATS_POLY_MOD > 0 && (e % (ATS_POLY_MOD > 0 ? ATS_POLY_MOD : 1)) == ATS_POLY_MOD - 1) {
            p0 = attn_exp2_poly3(x0);
            p1 = attn_exp2_poly3(x1);
        } else {
            p0 = ex2_approx(x0);
            p1 = ex2_approx(x1);
        }
        ls[e & 1] += p0 + p1;
        pk[e] = cvt_f16x2(p0, p1);
    }
#endif
}

// The same with the fp16 rounding of the MUFU pairs moved off the XU pipe (attention10.cuh, AT10_INTPACK).  F2FP runs on the
// XU pipe like MUFU.EX2, 8 cycles per warp instruction: one pack per two exponentials made it a third of the pipe that bounds
// the kernel.  Here probabilities are produced as P' = 2^7 P, and for a MUFU pair the exponent argument carries a further -112,
// i.e. MUFU returns v = P' 2^-112: the fp32 exponent field of v IS the fp16 exponent field of P' (127 - 112 = 15), so
//     fp16(P') = (bits(v) + 0x1000) >> 13          (round half up; fp32-denormal v would map onto fp16 denormals, .ftz drops them:
//                                                   P' < 2^-14, i.e. P < 2^-21 of the reference maximum, is zero)
// and two of them are packed with IMAD, IMAD, PRMT: hi16(bits * 8 + 0x8000).  The 2^7 puts the flush threshold at 2^-21 instead
// of 2^-14 (fp16 denormals reach 2^-24); lazily referenced probabilities stay <= 2^15 < 65504 as long as growth beyond 2^8 is
// re-exponentiated (attention10.cuh does).  Polynomial pairs are built in true scale (P', cubic + exponent patch as above) and
// still packed with F2FP: their clamp at 2^-30 needs the rounding to reach zero for masked keys.
// Row sums: `ls` accumulates the MUFU pairs in the v scale (multiply by 2^112 at the end), `lp` the polynomial pairs in P' scale.
constexpr float ATS_IP_SHIFT = 7.0f;               // P' = 2^7 P
constexpr float ATS_IP_BIAS = 112.0f - 7.0f;       // MUFU pairs: exponent argument offset so that v = P' 2^-112
constexpr float ATS_IP_UNBIAS = 5.192296858534828e33f;   // 2^112
// ATS_IP_MOD: only every ATS_IP_MOD-th pair is packed by integer arithmetic (1 = every MUFU pair), the others keep F2FP in true
// scale like the polynomial pairs — a way to balance the XU pipe (MUFU + F2FP) against the ALU pipe (LEA, LEA, PRMT).
#ifndef ATS_IP_MOD
#define ATS_IP_MOD 1
#endif
template <int E0, int E1>
__device__ __forceinline__ void attn_exp_pairs_ip(const uint32_t (&v)[32], uint32_t (&pk)[16], float c, float mc, float (&ls)[2], float (&lp)[2]) {
    const f32x2 c2 = pack_f32x2(c, c);
    const f32x2 nmc_m = pack_f32x2(-(mc + ATS_IP_BIAS), -(mc + ATS_IP_BIAS)), nmc_p = pack_f32x2(ATS_IP_SHIFT - mc, ATS_IP_SHIFT - mc);
    f32x2 acc_m = pack_f32x2(ls[0], ls[1]), acc_p = pack_f32x2(lp[0], lp[1]);
    float q0[E1 - E0], q1[E1 - E0];
    constexpr int PM = ATS_POLY_MOD > 0 ? ATS_POLY_MOD : 1;
#pragma unroll
    for (int e = E0; e < E1 + ATS_PIPE; ++e) {
        if (e < E1) {
            const bool poly = ATS_POLY_MOD > 0 && (e % PM) == PM - 1;
            const bool ip = !poly && (e % ATS_IP_MOD) == 0;          // integer pack (v scale); otherwise true scale + F2FP
            const f32x2 x = fma2_f32(pack_f32x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), c2, ip ? nmc_m : nmc_p);
            float p0, p1;
            if (poly) {
                float x0, x1;
                unpack_f32x2(x, x0, x1);
                const f32x2 t = pack_f32x2(fmaxf(x0, -30.0f), fmaxf(x1, -30.0f));
                const f32x2 magic = pack_f32x2(12582912.0f, 12582912.0f), nmagic = pack_f32x2(-12582912.0f, -12582912.0f);
                const f32x2 u = add2_f32(t, magic);
                const f32x2 w = add2_f32(u, nmagic);
                float w0, w1;
                unpack_f32x2(w, w0, w1);
                const f32x2 f = add2_f32(t, pack_f32x2(-w0, -w1));
                f32x2 q = fma2_f32(pack_f32x2(0.05508868396282196f, 0.05508868396282196f), f, pack_f32x2(0.24260404706001282f, 0.24260404706001282f));
                q = fma2_f32(q, f, pack_f32x2(0.6932762265205383f, 0.6932762265205383f));
                q = fma2_f32(q, f, pack_f32x2(0.9999289512634277f, 0.9999289512634277f));
                float r0, r1, u0, u1;
                unpack_f32x2(q, r0, r1);
                unpack_f32x2(u, u0, u1);
                p0 = __int_as_float(__float_as_int(r0) + (__float_as_int(u0) << 23));
                p1 = __int_as_float(__float_as_int(r1) + (__float_as_int(u1) << 23));
            } else {
                float x0, x1;
                unpack_f32x2(x, x0, x1);
                p0 = ex2_approx_ordered(x0);
                p1 = ex2_approx_ordered(x1);
            }
            q0[e - E0] = p0;
            q1[e - E0] = p1;
        }
        if (e - ATS_PIPE >= E0) {
            const int d = e - ATS_PIPE;
            const bool poly = ATS_POLY_MOD > 0 && (d % PM) == PM - 1;
            const bool ip = !poly && (d % ATS_IP_MOD) == 0;
            if (!ip) {
                acc_p = add2_f32_ordered(acc_p, pack_f32x2(q0[d - E0], q1[d - E0]));
                pk[d] = cvt_f16x2(q0[d - E0], q1[d - E0]);
            } else {
                acc_m = add2_f32_ordered(acc_m, pack_f32x2(q0[d - E0], q1[d - E0]));
                pk[d] = __byte_perm(__float_as_uint(q0[d - E0]) * 8u + 0x8000u, __float_as_uint(q1[d - E0]) * 8u + 0x8000u, 0x7632);
            }
        }
    }
    unpack_f32x2(acc_m, ls[0], ls[1]);
    unpack_f32x2(acc_p, lp[0], lp[1]);
}

}  // namespace dino
