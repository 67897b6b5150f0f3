// float_advance.h -- k sequential float32 additions `p = p + s` in closed form, bit for bit.
//
// The ray caster accumulates its sample positions exactly like the reference (`pos += step`, float32, round to nearest
// even; utils/tsdf/voxel_tsdf.cu:254-301), also through the samples it can prove empty.  While the accumulator stays
// inside one binade (fixed sign and exponent, ulp U) a step is an integer operation on its bit pattern:
//     p = m U,  s = (S + f) U  with S = floor(s / U),  0 <= f < 1:   fl(p + s) = (m + S + [f > 1/2]) U        for f != 1/2,
// independent of m; for f = 1/2 (a tie at every step) the result is the even neighbour, so after ONE step inside the
// binade m is even and from then on the increment is constant as well (S if S is even, S + 1 if it is odd).  Hence:
// after three real steps p1, p2, p3 in one binade, D = bits(p3) - bits(p2) is the increment of EVERY following step
// for as long as the result stays in the binade: mantissa field <= 0x7FFFFF when the magnitude grows (the exact sum is
// then below 2^(e+1), where the grid is still U -- and 2^(e+1) itself is representable on both grids), and >= 1 when
// it shrinks (the exact sum is then above 2^e; below it the grid would be U / 2).  Steps that cross a binade, zeros,
// denormals, infinities and NaN are simply taken for real.  tools/experiments/float_advance_check.cc checks this file against plain
// repeated addition on the CPU (random and adversarial operands).
// STATUS: verified but NOT used by the kernels.  Wired into raycast_kernel together with a coarse far-field distance
// map (skips of 64+ samples) it was bit-exact on the whole GPU parity suite and SLOWER (172.5 / 174.8 us per view against
// 162): a round costs ~90 instructions and the lanes of a warp rarely agree on taking it.  Kept as a building block for
// volumes whose free space is much larger than a room.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define TSDF_HD __host__ __device__ __forceinline__
#else
#define TSDF_HD inline
#endif

namespace tsdf {

TSDF_HD int float_bits(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int i; memcpy(&i, &f, 4); return i;
#endif
}
TSDF_HD float bits_float(int i) {
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f; memcpy(&f, &i, 4); return f;
#endif
}

constexpr int kAdvanceUnbounded = 0x7FFFFFFF;

// b1, b2, b3 = bit patterns of three consecutive accumulator values (two real steps apart).  Returns how many FURTHER
// steps may be taken as `bits += D` (0 = none: take real steps; kAdvanceUnbounded with D = 0 = the accumulator no
// longer moves).
TSDF_HD int advance_limit(int b1, int b2, int b3, int& D) {
  D = 0;
  if ((((b1 ^ b2) | (b2 ^ b3)) & (int)0xFF800000) != 0) return 0;  // not one binade
  const unsigned e = ((unsigned)b3 >> 23) & 0xFFu;
  if (e == 0u || e == 255u) return 0;                                 // zero / denormal / inf / NaN
  D = b3 - b2;
  if (D == 0) return kAdvanceUnbounded;
  const int mant = b3 & 0x7FFFFF;
  const int room = D > 0 ? 0x7FFFFF - mant : mant - 1;
  if (room <= 0) return 0;
  const int a = D > 0 ? D : -D;
#ifdef __CUDA_ARCH__
  // any n with n * a <= room is valid (a smaller n only leaves more steps to the next round), so the quotient is taken
  // in float -- no 32-bit integer division on the device -- a shade low, and checked
  int n = __float2int_rz(__fdividef((float)room, (float)a) * 0.99999f);
  if (n * a > room) --n;
  return n;
#else
  return room / a;
#endif
}

// reference form for one accumulator: the value after k additions of s
TSDF_HD float advance_exact(float p, float s, int k) {
  while (k > 0) {
    if (k < 6) { p = p + s; --k; continue; }
    const float p1 = p + s, p2 = p1 + s, p3 = p2 + s;
    k -= 3; p = p3;
    int D;
    int n = advance_limit(float_bits(p1), float_bits(p2), float_bits(p3), D);
    if (n > k) n = k;
    if (n > 0) { p = bits_float(float_bits(p3) + n * D); k -= n; }
  }
  return p;
}

// three accumulators advanced together (what a ray needs: the same k for x, y and z)
TSDF_HD void advance_exact3(float& x, float& y, float& z, float sx, float sy, float sz, int k) {
  while (k > 0) {
    if (k < 8) { x = x + sx; y = y + sy; z = z + sz; --k; continue; }
    const float x1 = x + sx, y1 = y + sy, z1 = z + sz;
    const float x2 = x1 + sx, y2 = y1 + sy, z2 = z1 + sz;
    x = x2 + sx; y = y2 + sy; z = z2 + sz;
    k -= 3;
    int dx, dy, dz;
    int n = k;
    const int nx = advance_limit(float_bits(x1), float_bits(x2), float_bits(x), dx);
    const int ny = advance_limit(float_bits(y1), float_bits(y2), float_bits(y), dy);
    const int nz = advance_limit(float_bits(z1), float_bits(z2), float_bits(z), dz);
    n = n < nx ? n : nx; n = n < ny ? n : ny; n = n < nz ? n : nz;
    if (n > 0) {
      x = bits_float(float_bits(x) + n * dx); y = bits_float(float_bits(y) + n * dy); z = bits_float(float_bits(z) + n * dz);
      k -= n;
    } else {  // some accumulator is at a binade edge (or near zero): a few plain steps, then try again
      for (int j = 0; j < 4 && k > 0; ++j, --k) { x = x + sx; y = y + sy; z = z + sz; }
    }
  }
}

}  // namespace tsdf
