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

// cos(2 pi i / 32), i = 0..8 (the rest by symmetry): enough for register radices up to 32.
constexpr float kCos32[9] = {1.0f,
                             0.98078528040323044913f,
                             0.92387953251128675613f,
                             0.83146961230254523708f,
                             0.70710678118654752440f,
                             0.55557023301960222474f,
                             0.38268343236508977173f,
                             0.19509032201612826785f,
                             0.0f};
SMFFT_CX float cos32(int i)
{
    i &= 31;
    return i <= 8 ? kCos32[i] : i <= 16 ? -kCos32[16 - i] : i <= 24 ? -kCos32[i - 16] : kCos32[32 - i];
}
SMFFT_CX float sin32(int i) { return cos32(i + 24); }  // sin(x) = cos(x - pi/2)

// a * exp(s * 2 pi i * NUM / DEN), s = -1 for DIR == 0 (forward), +1 for DIR == 1 (inverse).
// DEN divides 32.  Trivial rotations cost no multiplies.
template <int DIR, int NUM, int DEN>
SMFFT_DEV float2 mul_wconst(float2 a)
{
    static_assert(32 % DEN == 0, "constant twiddle modulus must divide 32");
    constexpr int i32 = (NUM * (32 / DEN)) & 31;
    if constexpr (i32 == 0) {
        return a;
    } else if constexpr (i32 == 8) {
        return DIR ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    } else if constexpr (i32 == 16) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (i32 == 24) {
        return DIR ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
    } else {
        constexpr float c = cos32(i32);
        constexpr float s = DIR ? sin32(i32) : -sin32(i32);
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

}  // namespace detail
}  // namespace smfft
