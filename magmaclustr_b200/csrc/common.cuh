// common.cuh -- shared device/host definitions of the B200 MagmaClustR hot path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/magma_b200.h"

#define MB_PI 3.14159265358979323846

// Flattened kernel descriptor passed by value to kernels (mb_kernel + HP offsets).
struct DevKern {
  int n_terms, n_prod, has_noise, n_hp;
  int types[MB_MAX_TERMS];
  int hp_off[MB_MAX_TERMS];
};

__host__ __device__ inline int mb_term_nhp(int type) {
  return (type == MB_KERN_SE || type == MB_KERN_LIN) ? 2 : 3;
}

inline DevKern make_devkern(const mb_kernel* k) {
  DevKern d;
  d.n_terms = k->n_terms;
  d.n_prod = k->n_prod;
  d.has_noise = k->has_noise;
  if (k->n_terms == 1 && k->types[0] == MB_KERN_CONV) {   // multi-output: n_prod = number of outputs
    for (int i = 0; i < MB_MAX_TERMS; ++i) { d.types[i] = i == 0 ? MB_KERN_CONV : 0; d.hp_off[i] = 0; }
    d.n_hp = 2 * k->n_prod + 1 + (k->has_noise ? k->n_prod : 0);
    return d;
  }
  int off = 0;
  for (int i = 0; i < MB_MAX_TERMS; ++i) {
    d.types[i] = (i < k->n_terms) ? k->types[i] : 0;
    d.hp_off[i] = off;
    if (i < k->n_terms) off += mb_term_nhp(k->types[i]);
  }
  d.n_hp = off + (k->has_noise ? 1 : 0);
  return d;
}

// Pairwise sums of the five native builders, /root/reference/src/code.cpp:6-131, accumulated
// over the input dimensions in order.
struct PairSums {
  double d2;  // cpp_dist:  sum (x - y)^2
  double l1;  // cpp_noise: sum |x - y|
  double xy;  // cpp_prod:  sum x * y
};

__device__ __forceinline__ PairSums pair_sums(const double* xa, const double* xb, int D) {
  PairSums s = {0.0, 0.0, 0.0};
  for (int c = 0; c < D; ++c) {
    double a = xa[c], b = xb[c];
    double diff = a - b;
    s.d2 += diff * diff;
    s.l1 += fabs(diff);
    s.xy += a * b;
  }
  return s;
}

// Exponentials of the HPs that every matrix element needs, computed once per thread:
// SE / RQ: (exp(v), exp(-l)[, -]); PERIO: (exp(v), exp(-l), pi / exp(period)); LIN: (exp(slope),
// exp(offset)); noise: exp(noise).  They are the very sub-expressions of kernels.R, so hoisting
// them does not change any rounding.
__device__ __forceinline__ void prep_hp(const DevKern& kd, const double* hp, double* hpe) {
  if (kd.types[0] == MB_KERN_CONV) {      // exp() of every entry: l = exp(l_t), S = exp(S_t), l_u = exp(l_u_t), exp(noise_o)
    for (int p = 0; p < kd.n_hp; ++p) hpe[p] = exp(hp[p]);
    return;
  }
  for (int t = 0; t < kd.n_terms; ++t) {
    const double* h = hp + kd.hp_off[t];
    double* e = hpe + kd.hp_off[t];
    const int type = kd.types[t];
    if (type == MB_KERN_LIN) { e[0] = exp(h[0]); e[1] = exp(h[1]); }
    else {
      e[0] = exp(h[0]);
      e[1] = exp(-h[1]);
      if (type == MB_KERN_PERIO) e[2] = MB_PI / exp(h[2]);
      if (type == MB_KERN_RQ) e[2] = h[2];
    }
  }
  if (kd.has_noise) hpe[kd.n_hp - 1] = exp(hp[kd.n_hp - 1]);
}

// One base kernel and, when WITH_DERIV, its derivatives w.r.t. its own HPs
// (/root/reference/R/kernels.R:164-398).  dv has mb_term_nhp(type) entries.
template <bool WITH_DERIV>
__device__ __forceinline__ double base_kernel(int type, const double* hp, const double* e,
                                              const PairSums& s, const double* xa,
                                              const double* xb, int D, double* dv) {
  if (type == MB_KERN_SE) {
    // kernels.R:182-189: top = exp(-l) * 0.5 * d2 ; K = exp(v - top)
    double top = e[1] * 0.5 * s.d2;
    double k = exp(hp[0] - top);
    if (WITH_DERIV) {
      dv[0] = k;
      // the reference evaluates exp(v) * top * exp(-top); k * top is the same number up to
      // one rounding and saves two exponentials per matrix element
      dv[1] = k * top;
    }
    return k;
  } else if (type == MB_KERN_PERIO) {
    // code.cpp:31-76 + kernels.R:243-263
    double w = e[2];
    double perio = 0.0, sderiv = 0.0;
    for (int c = 0; c < D; ++c) {
      double ang = w * fabs(xa[c] - xb[c]);
      double sn, cs;
      sincos(ang, &sn, &cs);
      perio += sn * sn;
      if (WITH_DERIV) sderiv += 2 * sn * cs * ang;
    }
    double el = e[1];
    double k = e[0] * exp(-2 * el * perio);
    if (WITH_DERIV) {
      dv[0] = k;
      dv[1] = e[0] * 2 * perio * el * exp(-2 * perio * el);
      dv[2] = k * 2 * el * sderiv;
    }
    return k;
  } else if (type == MB_KERN_RQ) {
    // kernels.R:312-333 (rq_scale raw)
    double el = e[1];
    double a = hp[2];
    double term = 1 + s.d2 * el / (2 * a);
    double k = e[0] * pow(term, -a);
    if (WITH_DERIV) {
      dv[0] = k;
      dv[1] = e[0] * s.d2 * 0.5 * el * pow(term, -a - 1);
      dv[2] = k * (s.d2 * el / (2 * a * term) - log(term));
    }
    return k;
  } else {
    // kernels.R:385-391: exp(slope) * prod + exp(offset)
    double k = e[0] * s.xy + e[1];
    if (WITH_DERIV) {
      dv[0] = e[0] * s.xy;
      dv[1] = e[1];
    }
    return k;
  }
}

// Covariance element of a compound kernel "(k1*k2*...) + k + ..." and (WITH_DERIV) all its
// HP derivatives, following the fold of kernel-to-matrices.R:163-340: the derivative w.r.t. an
// HP of a product factor is d(k_j) * prod_{i != j} k_i (left-to-right), of a summand d(k_j)
// alone; the noise term adds exp(noise) on coincident pairs (:440-447,:481-487).
// Multi-output convolution kernel with one latent process (/root/reference/R/kernels.R:50-140, vectorised branch) between
// two points whose LAST input column is the 0-based output index:
//   K = S_i S_j / (2 pi den)^(d/2) exp(-dist^2 / (2 den)),  den = l_i + l_j + l_u,  d = number of coordinates,
//   dK/dl_u_t = common l_u,  dK/dl_t_o = common (l_i [o_i = o] + l_j [o_j = o]),  dK/dS_t_o = K ([o_i = o] + [o_j = o]),
//   common = K (-d / (2 den) + dist^2 / (2 den^2));  noise: exp(noise_o) on coincident points of output o
//   (kernel-to-matrices.R:121-157).  HP layout: l_t[O], S_t[O], l_u_t, noise[O].
template <bool WITH_DERIV>
__device__ __forceinline__ double conv_eval(const DevKern& kd, const double* hpe, const double* xa, const double* xb, int D,
                                            bool with_noise, double* dval) {
  const int O = kd.n_prod, dc = D - 1;
  PairSums s = pair_sums(xa, xb, dc);
  const int oi = (int)xa[dc], oj = (int)xb[dc];
  const double li = hpe[oi], lj = hpe[oj], lu = hpe[2 * O];
  const double den = (li + lj) + lu;
  const double k = (hpe[O + oi] * hpe[O + oj]) / pow(2 * MB_PI * den, dc / 2.0) * exp(-0.5 * s.d2 / den);
  const bool same = (oi == oj) && s.l1 < 0.000001;
  if (WITH_DERIV) {
    for (int p = 0; p < kd.n_hp; ++p) dval[p] = 0.0;
    const double common = k * ((-0.5 * dc / den) + (0.5 * s.d2 / (den * den)));
    dval[2 * O] = common * lu;
    dval[oi] += common * li;
    dval[oj] += common * lj;
    dval[O + oi] += k;
    dval[O + oj] += k;
    if (kd.has_noise && same) dval[2 * O + 1 + oi] = hpe[2 * O + 1 + oi];
  }
  return (kd.has_noise && with_noise && same) ? k + hpe[2 * O + 1 + oi] : k;
}

template <bool WITH_DERIV>
__device__ __forceinline__ double cov_eval(const DevKern& kd, const double* hp, const double* hpe,
                                           const double* xa, const double* xb, int D,
                                           bool with_noise, double* dval) {
  if (kd.types[0] == MB_KERN_CONV) return conv_eval<WITH_DERIV>(kd, hpe, xa, xb, D, with_noise, dval);
  PairSums s = pair_sums(xa, xb, D);
  double val;
  if (kd.n_terms == 1) {
    val = base_kernel<WITH_DERIV>(kd.types[0], hp, hpe, s, xa, xb, D, dval);
  } else {
    double kv[MB_MAX_TERMS];
    double dv[MB_MAX_HP];
    for (int t = 0; t < kd.n_terms; ++t)
      kv[t] = base_kernel<WITH_DERIV>(kd.types[t], hp + kd.hp_off[t], hpe + kd.hp_off[t], s, xa, xb, D,
                                      dv + kd.hp_off[t]);
    val = 0.0;
    for (int t = 0; t < kd.n_prod; ++t) val = (t == 0) ? (0.0 + kv[0]) : val * kv[t];
    for (int t = kd.n_prod; t < kd.n_terms; ++t) val = val + kv[t];
    if (WITH_DERIV) {
      for (int t = 0; t < kd.n_terms; ++t) {
        int np = mb_term_nhp(kd.types[t]);
        for (int p = 0; p < np; ++p) {
          double out = dv[kd.hp_off[t] + p];
          if (t < kd.n_prod && kd.n_prod > 1) {
            out = 0.0;
            for (int u = 0; u < kd.n_prod; ++u) {
              double fct = (u == t) ? dv[kd.hp_off[t] + p] : kv[u];
              out = (u == 0) ? (0.0 + fct) : out * fct;
            }
          }
          dval[kd.hp_off[t] + p] = out;
        }
      }
    }
  }
  if (kd.has_noise) {
    double nz = (s.l1 < 0.000001) ? hpe[kd.n_hp - 1] : 0.0;
    if (with_noise) val += nz;
    if (WITH_DERIV) dval[kd.n_hp - 1] = nz;
  }
  return val;
}

#define MB_CUDA_OK(call)                                                         \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      mb_set_error(__FILE__, __LINE__, cudaGetErrorString(e_));                  \
      return -2;                                                                 \
    }                                                                            \
  } while (0)

void mb_set_error(const char* file, int line, const char* msg);
