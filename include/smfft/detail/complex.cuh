// smfft/detail/complex.cuh -- float2 complex helpers and compile-time twiddle constants.
#pragma once
#include <utility>

#include "platform.cuh"

namespace smfft {
namespace detail {

// compile-time loop: f(std::integral_constant<int, 0>{}) ... f(integral_constant<int, N-1>{})
template <class F, int... Is>
SMFFT_DEV void static_for_impl(F&& f, std::integer_sequence<int, Is...>)
{
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
SMFFT_DEV void static_for(F&& f)
{
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

SMFFT_DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SMFFT_DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SMFFT_DEV float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SMFFT_DEV float2 csqr(float2 a) { return make_float2(a.x * a.x - a.y * a.y, 2.0f * a.x * a.y); }

SMFFT_CX int ilog2_c(int n) { return n <= 1 ? 0 : 1 + ilog2_c(n >> 1); }
SMFFT_CX int brev_c(int v, int bits)
{
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// ---- two complex values side by side: the packed f32x2 ALU of sm_100 (FADD2 / FMUL2 / FFMA2) -----------------
// cpair holds the same element of TWO transforms in structure-of-arrays form, re = (re0, re1), im = (im0, im1):
// every complex operation is then a handful of packed instructions that serve both transforms, with no lane
// crossing (rotations by +-i are register renaming plus an operand negation, constants are broadcast immediates).
// Used by block_fft_dual.cuh.  The emulator build evaluates the same expressions lane by lane.
struct cpair {
    float2 re, im;
};
namespace pk {
#if !defined(SMFFT_EMU)
SMFFT_DEV float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
SMFFT_DEV float2 mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
SMFFT_DEV float2 fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
SMFFT_DEV float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SMFFT_DEV float2 mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
SMFFT_DEV float2 fma(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
SMFFT_DEV float2 neg(float2 a) { return make_float2(-a.x, -a.y); }  // folds into the consumer's operand modifier
SMFFT_DEV float2 sub(float2 a, float2 b) { return add(a, neg(b)); }
SMFFT_DEV float2 bc(float c) { return make_float2(c, c); }
}  // namespace pk
// one complex value, (re, im) in the two lanes: add / subtract as ONE packed instruction (BlockCfg::PACK kernels)
template <int PACK>
SMFFT_DEV float2 cadd_p(float2 a, float2 b)
{
    if constexpr (PACK) return pk::add(a, b); else return make_float2(a.x + b.x, a.y + b.y);
}
template <int PACK>
SMFFT_DEV float2 csub_p(float2 a, float2 b)
{
    if constexpr (PACK) return pk::sub(a, b); else return make_float2(a.x - b.x, a.y - b.y);
}
SMFFT_DEV cpair cadd(cpair a, cpair b) { return cpair{pk::add(a.re, b.re), pk::add(a.im, b.im)}; }
SMFFT_DEV cpair csub(cpair a, cpair b) { return cpair{pk::sub(a.re, b.re), pk::sub(a.im, b.im)}; }
SMFFT_DEV cpair cmul(cpair a, cpair b)
{
    return cpair{pk::fma(pk::neg(a.im), b.im, pk::mul(a.re, b.re)), pk::fma(a.re, b.im, pk::mul(a.im, b.re))};
}
SMFFT_DEV cpair csqr(cpair a)
{
    const float2 t = pk::mul(a.re, a.im);
    return cpair{pk::fma(pk::neg(a.im), a.im, pk::mul(a.re, a.re)), pk::add(t, t)};
}
// the same complex value in both lanes
SMFFT_DEV cpair cdup(float2 w) { return cpair{pk::bc(w.x), pk::bc(w.y)}; }
// value-type dispatch for code shared by the scalar and the packed paths
template <int PACK>
SMFFT_DEV cpair cadd_p(cpair a, cpair b) { return cadd(a, b); }
template <int PACK>
SMFFT_DEV cpair csub_p(cpair a, cpair b) { return csub(a, b); }
template <class V>
SMFFT_DEV V from_scalar(float2 w);
template <>
SMFFT_DEV float2 from_scalar<float2>(float2 w) { return w; }
template <>
SMFFT_DEV cpair from_scalar<cpair>(float2 w) { return cdup(w); }

// cos(2 pi i / 64), i = 0..16 (the rest by symmetry): register radices up to 32 and the W_{2R} constants
// of the real-transform pass need moduli up to 64.
constexpr float kCos64[17] = {1.0f,
                              0.99518472667219688624f,
                              0.98078528040323044913f,
                              0.95694033573220886494f,
                              0.92387953251128675613f,
                              0.88192126434835502971f,
                              0.83146961230254523708f,
                              0.77301045336273696081f,
                              0.70710678118654752440f,
                              0.63439328416364549822f,
                              0.55557023301960222474f,
                              0.47139673682599764856f,
                              0.38268343236508977173f,
                              0.29028467725446236764f,
                              0.19509032201612826785f,
                              0.09801714032956060199f,
                              0.0f};
SMFFT_CX float cos64(int i)
{
    i &= 63;
    return i <= 16 ? kCos64[i] : i <= 32 ? -kCos64[32 - i] : i <= 48 ? -kCos64[i - 32] : kCos64[64 - i];
}
SMFFT_CX float sin64(int i) { return cos64(i + 48); }  // sin(x) = cos(x - pi/2)

// a * exp(s * 2 pi i * NUM / DEN), s = -1 for DIR == 0 (forward), +1 for DIR == 1 (inverse).
// DEN divides 64.  Trivial rotations cost no multiplies.
template <int DIR, int NUM, int DEN>
SMFFT_DEV float2 mul_wconst(float2 a)
{
    static_assert(64 % DEN == 0, "constant twiddle modulus must divide 64");
    constexpr int i64 = (NUM * (64 / DEN)) & 63;
    if constexpr (i64 == 0) {
        return a;
    } else if constexpr (i64 == 16) {
        return DIR ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    } else if constexpr (i64 == 32) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (i64 == 48) {
        return DIR ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
    } else {
        constexpr float c = cos64(i64);
        constexpr float s = DIR ? sin64(i64) : -sin64(i64);
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

template <int DIR, int NUM, int DEN>
SMFFT_DEV cpair mul_wconst(cpair a)
{
    static_assert(64 % DEN == 0, "constant twiddle modulus must divide 64");
    constexpr int i64 = (NUM * (64 / DEN)) & 63;
    if constexpr (i64 == 0) {
        return a;
    } else if constexpr (i64 == 16) {
        return DIR ? cpair{pk::neg(a.im), a.re} : cpair{a.im, pk::neg(a.re)};
    } else if constexpr (i64 == 32) {
        return cpair{pk::neg(a.re), pk::neg(a.im)};
    } else if constexpr (i64 == 48) {
        return DIR ? cpair{a.im, pk::neg(a.re)} : cpair{pk::neg(a.im), a.re};
    } else {
        constexpr float c = cos64(i64);
        constexpr float s = DIR ? sin64(i64) : -sin64(i64);
        return cpair{pk::fma(a.re, pk::bc(c), pk::mul(a.im, pk::bc(-s))), pk::fma(a.im, pk::bc(c), pk::mul(a.re, pk::bc(s)))};
    }
}

}  // namespace detail
}  // namespace smfft
