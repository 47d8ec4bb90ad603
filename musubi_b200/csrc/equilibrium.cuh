// equilibrium.cuh -- equilibrium PDFs in the reference's sigma form, used by the ghost
// interpolation kernels (layout%quantities%pdfEq_ptr):
//   get_pdfEq_d3q19        mus_scheme_derived_quantities_type_module.f90:495-580
//   get_pdfEq_incomp_d3q19 ...:586-632
//   get_pdfEq_d3q27        ...:640-745
// Same operation order as the reference, so -fmad=false gives identical bits.
#pragma once
#include "common.cuh"

namespace musb200 {

__device__ __forceinline__ void sigmaD3Q19(double vx, double vy, double vz, double (&s)[23]) {
  s[22] = vx + vy; s[21] = vx - vy; s[20] = vx + vz; s[19] = vx - vz; s[18] = vy + vz; s[17] = vy - vz;
  s[16] = 3.0 * s[22]; s[15] = 3.0 * s[21]; s[14] = 3.0 * s[20];
  s[13] = 3.0 * s[19]; s[12] = 3.0 * s[18]; s[11] = 3.0 * s[17];
  s[10] = 4.5 * (s[22] * s[22]); s[9] = 4.5 * (s[21] * s[21]); s[8] = 4.5 * (s[20] * s[20]);
  s[7] = 4.5 * (s[19] * s[19]); s[6] = 4.5 * (s[18] * s[18]); s[5] = 4.5 * (s[17] * s[17]);
  s[4] = 4.5 * (vx * vx); s[3] = 4.5 * (vy * vy); s[2] = 4.5 * (vz * vz);
  s[1] = (1.0 / 3.0) * (s[2] + s[3] + s[4]);
}

__device__ __forceinline__ void pdfEqD3Q19(double rho, double vx, double vy, double vz,
                                           double (&f)[19]) {
  double s[23];
  sigmaD3Q19(vx, vy, vz, s);
  const double r18 = (1.0 / 18.0) * rho, r36 = (1.0 / 36.0) * rho;
  f[0] = -r18 * (3.0 * vx - s[4] + s[1] - 1.0);
  f[1] = -r18 * (3.0 * vy - s[3] + s[1] - 1.0);
  f[2] = -r18 * (3.0 * vz - s[2] + s[1] - 1.0);
  f[3] = r18 * (3.0 * vx + s[4] - s[1] + 1.0);
  f[4] = r18 * (3.0 * vy + s[3] - s[1] + 1.0);
  f[5] = r18 * (3.0 * vz + s[2] - s[1] + 1.0);
  f[6] = r36 * (s[6] - s[12] - s[1] + 1.0);
  f[7] = r36 * (s[5] - s[11] - s[1] + 1.0);
  f[8] = r36 * (s[5] + s[11] - s[1] + 1.0);
  f[9] = r36 * (s[6] + s[12] - s[1] + 1.0);
  f[10] = r36 * (s[8] - s[14] - s[1] + 1.0);
  f[11] = r36 * (s[7] + s[13] - s[1] + 1.0);
  f[12] = r36 * (s[7] - s[13] - s[1] + 1.0);
  f[13] = r36 * (s[8] + s[14] - s[1] + 1.0);
  f[14] = r36 * (s[10] - s[16] - s[1] + 1.0);
  f[15] = r36 * (s[9] - s[15] - s[1] + 1.0);
  f[16] = r36 * (s[9] + s[15] - s[1] + 1.0);
  f[17] = r36 * (s[10] + s[16] - s[1] + 1.0);
  f[18] = -(1.0 / 3.0) * rho * (s[1] - 1.0);
}

__device__ __forceinline__ void pdfEqIncompD3Q19(double rho, double vx, double vy, double vz,
                                                 double (&f)[19]) {
  double s[23];
  sigmaD3Q19(vx, vy, vz, s);
  const double rho0 = 1.0;
  const double r18 = (1.0 / 18.0) * rho, r36 = (1.0 / 36.0) * rho;
  const double z18 = (1.0 / 18.0) * rho0, z36 = (1.0 / 36.0) * rho0;
  f[0] = r18 - z18 * (3.0 * vx - s[4] + s[1]);
  f[1] = r18 - z18 * (3.0 * vy - s[3] + s[1]);
  f[2] = r18 - z18 * (3.0 * vz - s[2] + s[1]);
  f[3] = r18 + z18 * (3.0 * vx + s[4] - s[1]);
  f[4] = r18 + z18 * (3.0 * vy + s[3] - s[1]);
  f[5] = r18 + z18 * (3.0 * vz + s[2] - s[1]);
  f[6] = r36 + z36 * (s[6] - s[12] - s[1]);
  f[7] = r36 + z36 * (s[5] - s[11] - s[1]);
  f[8] = r36 + z36 * (s[5] + s[11] - s[1]);
  f[9] = r36 + z36 * (s[6] + s[12] - s[1]);
  f[10] = r36 + z36 * (s[8] - s[14] - s[1]);
  f[11] = r36 + z36 * (s[7] + s[13] - s[1]);
  f[12] = r36 + z36 * (s[7] - s[13] - s[1]);
  f[13] = r36 + z36 * (s[8] + s[14] - s[1]);
  f[14] = r36 + z36 * (s[10] - s[16] - s[1]);
  f[15] = r36 + z36 * (s[9] - s[15] - s[1]);
  f[16] = r36 + z36 * (s[9] + s[15] - s[1]);
  f[17] = r36 + z36 * (s[10] + s[16] - s[1]);
  f[18] = (1.0 / 3.0) * rho - (1.0 / 3.0) * rho0 * s[1];
}

__device__ __forceinline__ void pdfEqD3Q27(double rho, double vx, double vy, double vz,
                                           double (&f)[27]) {
  const double s34 = vx + vy, s33 = vx - vy, s32 = vx + vz, s31 = vx - vz;
  const double s30 = vy + vz, s29 = vy - vz;
  const double s28 = vx + vy + vz, s27 = vx + vy - vz, s26 = vx - vy + vz, s25 = vy - vx + vz;
  const double s24 = 3.0 * s34, s23 = 3.0 * s33, s22 = 3.0 * s32, s21 = 3.0 * s31;
  const double s20 = 3.0 * s30, s19 = 3.0 * s29, s18 = 3.0 * s28, s17 = 3.0 * s27;
  const double s16 = 3.0 * s26, s15 = 3.0 * s25;
  const double s14 = 4.5 * (s34 * s34), s13 = 4.5 * (s33 * s33), s12 = 4.5 * (s32 * s32);
  const double s11 = 4.5 * (s31 * s31), s10 = 4.5 * (s30 * s30), s9 = 4.5 * (s29 * s29);
  const double s8 = 4.5 * (s28 * s28), s7 = 4.5 * (s27 * s27), s6 = 4.5 * (s26 * s26);
  const double s5 = 4.5 * (s25 * s25);
  const double s4 = 4.5 * (vx * vx), s3 = 4.5 * (vy * vy), s2 = 4.5 * (vz * vz);
  const double s1 = (1.0 / 3.0) * (s2 + s3 + s4);
  const double r27 = (2.0 / 27.0) * rho, r54 = (1.0 / 54.0) * rho, r216 = (1.0 / 216.0) * rho;
  f[0] = -r27 * (3.0 * vx - s4 + s1 - 1.0);
  f[1] = -r27 * (3.0 * vy - s3 + s1 - 1.0);
  f[2] = -r27 * (3.0 * vz - s2 + s1 - 1.0);
  f[3] = r27 * (3.0 * vx + s4 - s1 + 1.0);
  f[4] = r27 * (3.0 * vy + s3 - s1 + 1.0);
  f[5] = r27 * (3.0 * vz + s2 - s1 + 1.0);
  f[6] = r54 * (s10 - s20 - s1 + 1.0);
  f[7] = r54 * (s9 - s19 - s1 + 1.0);
  f[8] = r54 * (s9 + s19 - s1 + 1.0);
  f[9] = r54 * (s10 + s20 - s1 + 1.0);
  f[10] = r54 * (s12 - s22 - s1 + 1.0);
  f[11] = r54 * (s11 + s21 - s1 + 1.0);
  f[12] = r54 * (s11 - s21 - s1 + 1.0);
  f[13] = r54 * (s12 + s22 - s1 + 1.0);
  f[14] = r54 * (s14 - s24 - s1 + 1.0);
  f[15] = r54 * (s13 - s23 - s1 + 1.0);
  f[16] = r54 * (s13 + s23 - s1 + 1.0);
  f[17] = r54 * (s14 + s24 - s1 + 1.0);
  f[18] = -r216 * (s18 - s8 + s1 - 1.0);
  f[19] = -r216 * (s17 - s7 + s1 - 1.0);
  f[20] = -r216 * (s16 - s6 + s1 - 1.0);
  f[21] = r216 * (s15 + s5 - s1 + 1.0);
  f[22] = -r216 * (s15 - s5 + s1 - 1.0);
  f[23] = r216 * (s16 + s6 - s1 + 1.0);
  f[24] = r216 * (s17 + s7 - s1 + 1.0);
  f[25] = r216 * (s18 + s8 - s1 + 1.0);
  f[26] = -(8.0 / 27.0) * rho * (s1 - 1.0);
}

// get_pdfEq_incomp_d3q27 (...:751-811)
__device__ __forceinline__ void pdfEqIncompD3Q27(double rho, double vx, double vy, double vz,
                                           double (&f)[27]) {
  const double s34 = vx + vy, s33 = vx - vy, s32 = vx + vz, s31 = vx - vz;
  const double s30 = vy + vz, s29 = vy - vz;
  const double s28 = vx + vy + vz, s27 = vx + vy - vz, s26 = vx - vy + vz, s25 = vy - vx + vz;
  const double s24 = 3.0 * s34, s23 = 3.0 * s33, s22 = 3.0 * s32, s21 = 3.0 * s31;
  const double s20 = 3.0 * s30, s19 = 3.0 * s29, s18 = 3.0 * s28, s17 = 3.0 * s27;
  const double s16 = 3.0 * s26, s15 = 3.0 * s25;
  const double s14 = 4.5 * (s34 * s34), s13 = 4.5 * (s33 * s33), s12 = 4.5 * (s32 * s32);
  const double s11 = 4.5 * (s31 * s31), s10 = 4.5 * (s30 * s30), s9 = 4.5 * (s29 * s29);
  const double s8 = 4.5 * (s28 * s28), s7 = 4.5 * (s27 * s27), s6 = 4.5 * (s26 * s26);
  const double s5 = 4.5 * (s25 * s25);
  const double s4 = 4.5 * (vx * vx), s3 = 4.5 * (vy * vy), s2 = 4.5 * (vz * vz);
  const double s1 = (1.0 / 3.0) * (s2 + s3 + s4);
  const double r27 = (2.0 / 27.0) * rho, r54 = (1.0 / 54.0) * rho, r216 = (1.0 / 216.0) * rho;
  const double rho0 = 1.0;
  const double z27 = (2.0 / 27.0) * rho0, z54 = (1.0 / 54.0) * rho0, z216 = (1.0 / 216.0) * rho0;
  f[0] = r27 - z27 * (3.0 * vx - s4 + s1);
  f[1] = r27 - z27 * (3.0 * vy - s3 + s1);
  f[2] = r27 - z27 * (3.0 * vz - s2 + s1);
  f[3] = r27 + z27 * (3.0 * vx + s4 - s1);
  f[4] = r27 + z27 * (3.0 * vy + s3 - s1);
  f[5] = r27 + z27 * (3.0 * vz + s2 - s1);
  f[6] = r54 + z54 * (s10 - s20 - s1);
  f[7] = r54 + z54 * (s9 - s19 - s1);
  f[8] = r54 + z54 * (s9 + s19 - s1);
  f[9] = r54 + z54 * (s10 + s20 - s1);
  f[10] = r54 + z54 * (s12 - s22 - s1);
  f[11] = r54 + z54 * (s11 + s21 - s1);
  f[12] = r54 + z54 * (s11 - s21 - s1);
  f[13] = r54 + z54 * (s12 + s22 - s1);
  f[14] = r54 + z54 * (s14 - s24 - s1);
  f[15] = r54 + z54 * (s13 - s23 - s1);
  f[16] = r54 + z54 * (s13 + s23 - s1);
  f[17] = r54 + z54 * (s14 + s24 - s1);
  f[18] = r216 - z216 * (s18 - s8 + s1);
  f[19] = r216 - z216 * (s17 - s7 + s1);
  f[20] = r216 - z216 * (s16 - s6 + s1);
  f[21] = r216 + z216 * (s15 + s5 - s1);
  f[22] = r216 - z216 * (s15 - s5 + s1);
  f[23] = r216 + z216 * (s16 + s6 - s1);
  f[24] = r216 + z216 * (s17 + s7 - s1);
  f[25] = r216 + z216 * (s18 + s8 - s1);
  f[26] = (8.0 / 27.0) * rho - (8.0 / 27.0) * rho0 * s1;
}

}  // namespace musb200
