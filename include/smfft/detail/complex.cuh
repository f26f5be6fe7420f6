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

#if defined(SMFFT_PACKED_F32X2) && !defined(SMFFT_EMU)
SMFFT_DEV float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
SMFFT_DEV float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
#else
SMFFT_DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SMFFT_DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
SMFFT_DEV float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SMFFT_DEV float2 csqr(float2 a) { return make_float2(a.x * a.x - a.y * a.y, 2.0f * a.x * a.y); }

SMFFT_CX int ilog2_c(int n) { return n <= 1 ? 0 : 1 + ilog2_c(n >> 1); }
SMFFT_CX int brev_c(int v, int bits)
{
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

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

}  // namespace detail
}  // namespace smfft
