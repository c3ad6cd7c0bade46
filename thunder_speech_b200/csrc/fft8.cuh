// In-register 8-point forward DFT (W = exp(-2*pi*i/8)), natural-order input and output.
// Decimation in frequency: one radix-2 split, two 4-point DFTs.  Compiled for host too so that the
// CPU test-suite can check it against a naive DFT (tests/test_host_units.py via csrc/host_selftest.cpp).
#pragma once
#if defined(__CUDACC__)
#define TS_HD __host__ __device__ __forceinline__
#else
#define TS_HD inline
#endif

namespace ts {

TS_HD void dft8(float (&r)[8], float (&i)[8]) {
  const float h = 0.70710678118654752440f;
  // radix-2 split
  const float a0r = r[0] + r[4], a0i = i[0] + i[4];
  const float a4r = r[0] - r[4], a4i = i[0] - i[4];
  const float a1r = r[1] + r[5], a1i = i[1] + i[5];
  const float b5r = r[1] - r[5], b5i = i[1] - i[5];
  const float a2r = r[2] + r[6], a2i = i[2] + i[6];
  const float b6r = r[2] - r[6], b6i = i[2] - i[6];
  const float a3r = r[3] + r[7], a3i = i[3] + i[7];
  const float b7r = r[3] - r[7], b7i = i[3] - i[7];
  // odd-branch twiddles W8^1 = (1-i)/sqrt2, W8^2 = -i, W8^3 = (-1-i)/sqrt2
  const float a5r = (b5r + b5i) * h, a5i = (b5i - b5r) * h;
  const float a6r = b6i, a6i = -b6r;
  const float a7r = (b7i - b7r) * h, a7i = -(b7r + b7i) * h;
  // even outputs: DFT4(a0, a1, a2, a3)
  {
    const float c0r = a0r + a2r, c0i = a0i + a2i;
    const float c1r = a0r - a2r, c1i = a0i - a2i;
    const float c2r = a1r + a3r, c2i = a1i + a3i;
    const float c3r = a1i - a3i, c3i = -(a1r - a3r);  // (a1 - a3) * (-i)
    r[0] = c0r + c2r; i[0] = c0i + c2i;
    r[4] = c0r - c2r; i[4] = c0i - c2i;
    r[2] = c1r + c3r; i[2] = c1i + c3i;
    r[6] = c1r - c3r; i[6] = c1i - c3i;
  }
  // odd outputs: DFT4(a4, a5, a6, a7)
  {
    const float d0r = a4r + a6r, d0i = a4i + a6i;
    const float d1r = a4r - a6r, d1i = a4i - a6i;
    const float d2r = a5r + a7r, d2i = a5i + a7i;
    const float d3r = a5i - a7i, d3i = -(a5r - a7r);
    r[1] = d0r + d2r; i[1] = d0i + d2i;
    r[5] = d0r - d2r; i[5] = d0i - d2i;
    r[3] = d1r + d3r; i[3] = d1i + d3i;
    r[7] = d1r - d3r; i[7] = d1i - d3i;
  }
}

}  // namespace ts
