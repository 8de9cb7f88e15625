// ookd_common.cuh -- shared device/host definitions for the sm_100a receive path.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "ookd_gpu.h"

namespace ookd {

typedef unsigned long long u64;
typedef long long i64;

// ---------------------------------------------------------------------------------------
// Exact arithmetic helpers.  The reference accumulates  acc = acc + taps[i] * x[n-i]  with a
// rounded multiply followed by a rounded add (src/fir.c:313-318, x86-64 baseline has no FMA).
// __fmul_rn / __fadd_rn are never contracted by nvcc/ptxas.  (Do NOT use the packed
// mul.rn.f32x2 / add.rn.f32x2 forms here: ptxas 12.9 fuses that pair into one FFMA2 even with
// explicit .rn -- verified in tools/ubench_fp.cu's SASS -- which changes the rounding.)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float mac_exact(float acc, float tap, float x)
{
    return __fadd_rn(acc, __fmul_rn(tap, x));
}

// SC16Q11 -> float, src/complexf.h:68-77:  (float) k * (1.0f / 2048.0f).  Both the int->float
// conversion (|k| < 2^15) and the scaling by 2^-11 are exact, so any exact route is identical.
// Route used here: place (k + 32768) in the mantissa of a float in [2^12, 2^13) (ulp 2^-11),
// which reads as 4096 + (k + 32768)/2048, then subtract 4112.  One LOP3 + one PRMT + one FADD
// per component instead of an I2F (quarter-rate conversion pipe) + FMUL.
__device__ __forceinline__ float2 sc16q11_to_float2(uint32_t w)
{
    const uint32_t u = w ^ 0x80008000u;                       // offset-binary halves
    const uint32_t bi = __byte_perm(u, 0x45800000u, 0x7610);  // [0x45,0x80,u.b1,u.b0]
    const uint32_t bq = __byte_perm(u, 0x45800000u, 0x7632);  // [0x45,0x80,u.b3,u.b2]
    return make_float2(__fadd_rn(__uint_as_float(bi), -4112.0f),
                       __fadd_rn(__uint_as_float(bq), -4112.0f));
}

// power as the reference computes it (src/complexf.h:43-46): re*re + im*im, two rounded
// multiplies and one rounded add.
__device__ __forceinline__ float power_exact(float re, float im)
{
    return __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
}

// ---------------------------------------------------------------------------------------
// Compiled state machine as it lives in device memory (see sm_compile.c).
// ---------------------------------------------------------------------------------------
#define OOKD_SM_MAX_STATES   64
#define OOKD_SM_MAX_TRIGGERS 256

struct SmTable {
    uint32_t num_states, num_triggers, max_bits, k_sat;
    ookd_sm_state_k   states[OOKD_SM_MAX_STATES];
    ookd_sm_trigger_k triggers[OOKD_SM_MAX_TRIGGERS];
};

// State between two output samples (device twin of ookd_sm_carry; data as 4 x u64).
struct SmCarry {
    uint32_t state, k, num_bits, prev;
    u64 data[4];
};

struct SmMsg {
    u64 out_sample;
    uint32_t num_bits, pad;
    u64 data[4];
};

}  // namespace ookd
