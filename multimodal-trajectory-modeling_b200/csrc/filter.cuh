// Per-(trajectory, cluster) marginal log-likelihood of a linear-Gaussian state-space model in
// prediction-error (sequential conditioning) form, register resident, fp64.
//
// Replaces the arithmetic of the reference's
//   statespace.mm / CC  (framework/marginalizable_state_space_model.py:259-364) and
//   multivariate_normal_log_likelihood (:728-773)
// without ever forming the dense (d+l)T covariance: the reference itself checks that the dense
// joint density equals the step-wise generative one (:993-1002).
//
// Row-vector convention (:8-16):  z_1 ~ N(m, S),  z_{t+1} = z_t A + N(0, Gamma),  x_t = z_t H + N(0, Lambda).
//
// For each time step t the thread holds the Gaussian of z_t given everything already conditioned
// on, (mu, P).  Present coordinates of z_t are conditioned on one scalar at a time (an LDL'
// factorisation of the present sub-block, one pivot per scalar); then the predictive of x_t,
// (mu H, P H, H' P H + Lambda), is formed and its present coordinates are conditioned on the same
// way; then (mu, P) <- (mu A, A' P A + Gamma).  Each pivot s with innovation v contributes
// v^2 / s to the quadratic form and s to the determinant; missing coordinates contribute nothing
// (gain 0), which is exactly the reference's sub-setting of mean and covariance.
//
// A coordinate counts as present iff (y - mean) is finite with `mean` the UNCONDITIONAL stacked
// mean of the reference (:769-770) — so +-inf data and coordinates whose model mean is NaN are
// both "missing".  The per-(cluster, t) finiteness of that mean is precomputed as a bit mask.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mtm_b200.h"

namespace mtm {

// Compile-time (D_, L_ > 0: loops unrolled, state in registers) or run-time (D_ == 0: generic
// fallback for any d, l <= MTM_MAX_DIM, state in local memory) dimensions behind one source.
template <int D_, int L_>
struct Dims {
    static constexpr bool kStatic = (D_ > 0);
    static constexpr int DM = kStatic ? D_ : MTM_MAX_DIM;
    static constexpr int LM = kStatic ? L_ : MTM_MAX_DIM;
    int d_rt, l_rt;
    __host__ __device__ __forceinline__ int d() const { return kStatic ? D_ : d_rt; }
    __host__ __device__ __forceinline__ int l() const { return kStatic ? L_ : l_rt; }
    __host__ __device__ __forceinline__ int stride() const {
        return 1 + d() + 3 * d() * d() + d() * l() + l() * l();
    }
};

#define MTM_UNROLL _Pragma("unroll")

__device__ __forceinline__ bool finite_bits(double y) {
    // exponent field != all ones; integer pipe, keeps the FP64 pipe free
    return (__double2hiint(y) & 0x7ff00000) != 0x7ff00000;
}

// 1/s to ~1 ulp: hardware seed r0 (MUFU.RCP64H, ~2^-20) + ONE third-order step
//   e = 1 - s r0,   1/s = r0 (1 + e + e^2 + e^3 + ...)  ->  r = r0 + r0 (e + e^2),  error e^3 ~ 2^-60
// = 3 dependent DFMAs (the IEEE division is ~10 instructions plus a slow path, two Newton steps are 4).
// s = 0 -> NaN, s = +-inf -> NaN, NaN -> NaN: all of these are flagged by the caller.
__device__ __forceinline__ double fast_rcp(double s) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    const double e = fma(-s, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// Where the packed parameter rows [pi | m | S | A | G | H | Lam] live.
//  * ConstSrc: __constant__ memory, indexed with a warp-uniform offset.  The compiler fetches such
//    operands through the uniform datapath (ULDC) and feeds them to DFMA as uniform-register
//    operands, so a DFMA that multiplies by a model parameter reads only TWO vector registers and
//    issues at the full FP64 rate (tools/dfma_probe.cu: 34-35.7 TFLOP/s vs 24.6 with three vector
//    sources, 27-31 when the parameter first goes through a vector register via shared memory).
//  * SmemSrc: shared memory (fallback when n_c * stride exceeds the constant bank, and for the
//    generic run-time-dimension instantiation).
#ifndef MTM_CONST_DOUBLES
#define MTM_CONST_DOUBLES 7680  // 60 KB of the 64 KB constant bank
#endif
#ifdef MTM_DEFINE_CONST_PARAMS
__constant__ double c_par[MTM_CONST_DOUBLES];
struct ConstSrc {
    int base;
    __device__ __forceinline__ double operator[](int k) const { return c_par[base + k]; }
    __device__ __forceinline__ ConstSrc shifted(int o) const { return ConstSrc{base + o}; }
};
#endif
struct SmemSrc {
    const double *row;
    __device__ __forceinline__ double operator[](int k) const { return row[k]; }
    __device__ __forceinline__ SmemSrc shifted(int o) const { return SmemSrc{row + o}; }
};

template <class DM, class PS>
struct ParamView {
    PS ps;
    int oS, oA, oG, oH, oL;
    __device__ __forceinline__ ParamView(const DM &dm, const PS &src) : ps(src) {
        const int d = dm.d(), l = dm.l();
        oS = 1 + d;
        oA = oS + d * d;
        oG = oA + d * d;
        oH = oG + d * d;
        oL = oH + d * l;
    }
    __device__ __forceinline__ double pi() const { return ps[0]; }
    __device__ __forceinline__ double m(int k) const { return ps[1 + k]; }
    __device__ __forceinline__ double S(int k) const { return ps[oS + k]; }
    __device__ __forceinline__ double A(int k) const { return ps[oA + k]; }
    __device__ __forceinline__ double G(int k) const { return ps[oG + k]; }
    __device__ __forceinline__ double H(int k) const { return ps[oH + k]; }
    __device__ __forceinline__ double Lam(int k) const { return ps[oL + k]; }
};

template <class DM>
struct FilterState {
    double mu[DM::DM];
    double P[DM::DM][DM::DM];  // symmetric; only i <= j is kept
    double q;                  // sum v^2 / s
    double mant;               // running product of pivots, mantissa part (sign kept)
    int expo;                  // ... and its binary exponent
    int k;                     // number of present scalars
};

// element (i, j) of a symmetric matrix stored in its upper triangle
template <class M>
__device__ __forceinline__ double symget(const M &P, int i, int j) {
    return i <= j ? P[i][j] : P[j][i];
}

template <class DM, class PV>
__device__ __forceinline__ void filter_init(const DM &dm, FilterState<DM> &f, const PV &p) {
    const int d = dm.d();
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            f.mu[i] = p.m(i);
            MTM_UNROLL
            for (int j = i; j < DM::DM; ++j)
                if (j < d) f.P[i][j] = p.S(i * d + j);
        }
    }
    f.q = 0.0;
    f.mant = 1.0;
    f.expo = 0;
    f.k = 0;
}

// (mu, P) <- (mu A, A' P A + Gamma)     [SSM:62-64 for P, :312-314 for mu]
//
// All small products below are written in "outer-product order": the innermost loop runs over
// independent accumulators that share ONE multiplicand.  Consecutive DFMAs then (a) are independent
// (no latency chain) and (b) keep that multiplicand in the operand-reuse cache, which matters on
// this part: a DFMA reading three distinct vector registers issues at 72 % of the FP64 peak, one
// with a reused operand at 92 % (tools/dfma_probe.cu).
template <class DM, class PV>
__device__ __forceinline__ void filter_propagate(const DM &dm, FilterState<DM> &f, const PV &p) {
    const int d = dm.d();
    double nmu[DM::DM];
    double PA[DM::DM][DM::DM];
    MTM_UNROLL
    for (int j = 0; j < DM::DM; ++j) nmu[j] = 0.0;
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            const double mi = f.mu[i];
            MTM_UNROLL
            for (int j = 0; j < DM::DM; ++j)
                if (j < d) nmu[j] = fma(mi, p.A(i * d + j), nmu[j]);
        }
    }
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            MTM_UNROLL
            for (int j = 0; j < DM::DM; ++j) PA[i][j] = 0.0;
            MTM_UNROLL
            for (int a = 0; a < DM::DM; ++a) {
                if (a < d) {
                    const double pia = symget(f.P, i, a);
                    MTM_UNROLL
                    for (int j = 0; j < DM::DM; ++j)
                        if (j < d) PA[i][j] = fma(pia, p.A(a * d + j), PA[i][j]);
                }
            }
        }
    }
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            f.mu[i] = nmu[i];
            MTM_UNROLL
            for (int j = i; j < DM::DM; ++j)
                if (j < d) f.P[i][j] = p.G(i * d + j);
            MTM_UNROLL
            for (int a = 0; a < DM::DM; ++a) {
                if (a < d) {
                    const double aai = p.A(a * d + i);
                    MTM_UNROLL
                    for (int j = i; j < DM::DM; ++j)
                        if (j < d) f.P[i][j] = fma(aai, PA[a][j], f.P[i][j]);
                }
            }
        }
    }
}

// condition on the present coordinates of z_t (noise-free observations of state coordinates)
template <class DM>
__device__ __forceinline__ void filter_condition_z(const DM &dm, FilterState<DM> &f, const double (&zt)[DM::DM],
                                                   uint32_t finmask) {
    const int d = dm.d();
    MTM_UNROLL
    for (int j = 0; j < DM::DM; ++j) {
        if (j < d) {
            const double y = zt[j];
            const bool obs = finite_bits(y) && ((finmask >> j) & 1u);
            const double s = f.P[j][j];
            const double r = obs ? fast_rcp(s) : 0.0;
            const double v = obs ? (y - f.mu[j]) : 0.0;
            f.q = fma(v * v, r, f.q);
            if (obs) f.mant *= s;  // predicated DMUL (a select against 1.0 costs four IMAD moves, which steal FP64 issue slots)
            f.k += obs ? 1 : 0;
            double col[DM::DM];
            MTM_UNROLL
            for (int i = 0; i < DM::DM; ++i)
                if (i < d) col[i] = symget(f.P, i, j);
            // Row / column j of P and mu[j] are not updated here: when the coordinate is present they are
            // replaced below by exact values (0 and y), when it is missing the gain is 0 anyway.
            MTM_UNROLL
            for (int i = 0; i < DM::DM; ++i) {
                if (i < d && i != j) {
                    const double g = col[i] * r;
                    f.mu[i] = fma(g, v, f.mu[i]);
                    MTM_UNROLL
                    for (int b = i; b < DM::DM; ++b)
                        if (b < d && b != j) f.P[i][b] = fma(-g, col[b], f.P[i][b]);
                }
            }
            // the conditioned coordinate is now known exactly
            MTM_UNROLL
            for (int i = 0; i < DM::DM; ++i) {
                if (i < d) {
                    if (i <= j) f.P[i][j] = obs ? 0.0 : f.P[i][j];
                    else f.P[j][i] = obs ? 0.0 : f.P[j][i];
                }
            }
            f.mu[j] = obs ? y : f.mu[j];
        }
    }
}

// predictive of x_t and conditioning on its present coordinates   [SSM:161-166, 217-222, 337-339]
// LAST = true (final time step): the posterior of z_t is never used again, so the updates of
// (mu, P) are skipped; pivots, innovations and the x-block updates are unchanged.
template <class DM, class PV>
__device__ __forceinline__ void filter_condition_x(const DM &dm, FilterState<DM> &f, const PV &p,
                                                   const double (&xt)[DM::LM], uint32_t finmask, const bool LAST) {
    const int d = dm.d(), l = dm.l();
    double mux[DM::LM];
    double C[DM::DM][DM::LM];   // Cov(z_t, x_t | past) = P H
    double Sx[DM::LM][DM::LM];  // Var(x_t | past) = H' P H + Lambda, upper triangle
    MTM_UNROLL
    for (int b = 0; b < DM::LM; ++b) mux[b] = 0.0;
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            const double mi = f.mu[i];
            MTM_UNROLL
            for (int b = 0; b < DM::LM; ++b)
                if (b < l) mux[b] = fma(mi, p.H(i * l + b), mux[b]);
        }
    }
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            MTM_UNROLL
            for (int b = 0; b < DM::LM; ++b) C[i][b] = 0.0;
            MTM_UNROLL
            for (int a = 0; a < DM::DM; ++a) {
                if (a < d) {
                    const double pia = symget(f.P, i, a);
                    MTM_UNROLL
                    for (int b = 0; b < DM::LM; ++b)
                        if (b < l) C[i][b] = fma(pia, p.H(a * l + b), C[i][b]);
                }
            }
        }
    }
    MTM_UNROLL
    for (int a = 0; a < DM::LM; ++a) {
        if (a < l) {
            MTM_UNROLL
            for (int b = a; b < DM::LM; ++b)
                if (b < l) Sx[a][b] = p.Lam(a * l + b);
            MTM_UNROLL
            for (int i = 0; i < DM::DM; ++i) {
                if (i < d) {
                    const double hia = p.H(i * l + a);
                    MTM_UNROLL
                    for (int b = a; b < DM::LM; ++b)
                        if (b < l) Sx[a][b] = fma(hia, C[i][b], Sx[a][b]);
                }
            }
        }
    }
    MTM_UNROLL
    for (int j = 0; j < DM::LM; ++j) {
        if (j < l) {
            const double y = xt[j];
            const bool obs = finite_bits(y) && ((finmask >> j) & 1u);
            const double s = Sx[j][j];
            const double r = obs ? fast_rcp(s) : 0.0;
            const double v = obs ? (y - mux[j]) : 0.0;
            f.q = fma(v * v, r, f.q);
            if (obs) f.mant *= s;  // predicated DMUL (a select against 1.0 costs four IMAD moves, which steal FP64 issue slots)
            f.k += obs ? 1 : 0;
            MTM_UNROLL
            for (int i = 0; i < DM::DM; ++i) {
                if (i < d) {
                    const double g = C[i][j] * r;
                    if (!LAST) {
                        f.mu[i] = fma(g, v, f.mu[i]);
                        MTM_UNROLL
                        for (int a = i; a < DM::DM; ++a)
                            if (a < d) f.P[i][a] = fma(-g, C[a][j], f.P[i][a]);
                    }
                    MTM_UNROLL
                    for (int b = j + 1; b < DM::LM; ++b)
                        if (b < l) C[i][b] = fma(-g, Sx[j][b], C[i][b]);
                }
            }
            MTM_UNROLL
            for (int a = j + 1; a < DM::LM; ++a) {
                if (a < l) {
                    const double g = Sx[j][a] * r;
                    mux[a] = fma(g, v, mux[a]);
                    MTM_UNROLL
                    for (int b = a; b < DM::LM; ++b)
                        if (b < l) Sx[a][b] = fma(-g, Sx[j][b], Sx[a][b]);
                }
            }
        }
    }
}

// ---- point-mass shortcuts ------------------------------------------------------------------------------------
// When EVERY hidden-state coordinate of a step is present, the posterior of z_t is the point mass at z_t:
// P is exactly 0 (filter_condition_z zeroes the rows) and mu == z_t.  The generic code would then only move
// zeros around (C = P H = 0, gains 0, A'PA = 0); these two functions skip that work and give bit-identical
// numbers.  They are taken on a WARP-UNIFORM vote (all 32 trajectories of the warp), so there is no
// divergence; storing the resident data sorted by its per-step "z complete" pattern (engine.py) makes the vote
// succeed for ~41 % of the steps at 20 % element-wise missingness, saving ~26 % of the FP64 work.

// (mu, P) <- (mu A, Gamma): the propagation of a point mass
template <class DM, class PV>
__device__ __forceinline__ void filter_propagate_point(const DM &dm, FilterState<DM> &f, const PV &p) {
    const int d = dm.d();
    double nmu[DM::DM];
    MTM_UNROLL
    for (int j = 0; j < DM::DM; ++j) nmu[j] = 0.0;
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            const double mi = f.mu[i];
            MTM_UNROLL
            for (int j = 0; j < DM::DM; ++j)
                if (j < d) nmu[j] = fma(mi, p.A(i * d + j), nmu[j]);
        }
    }
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            f.mu[i] = nmu[i];
            MTM_UNROLL
            for (int j = i; j < DM::DM; ++j)
                if (j < d) f.P[i][j] = p.G(i * d + j);
        }
    }
}

// x_t given a known z_t: mean z_t H, covariance Lambda; only the x-block is conditioned, (mu, P) stay.
template <class DM, class PV>
__device__ __forceinline__ void filter_condition_x_point(const DM &dm, FilterState<DM> &f, const PV &p,
                                                         const double (&xt)[DM::LM], uint32_t finmask) {
    const int d = dm.d(), l = dm.l();
    double mux[DM::LM];
    double Sx[DM::LM][DM::LM];
    MTM_UNROLL
    for (int b = 0; b < DM::LM; ++b) mux[b] = 0.0;
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            const double mi = f.mu[i];
            MTM_UNROLL
            for (int b = 0; b < DM::LM; ++b)
                if (b < l) mux[b] = fma(mi, p.H(i * l + b), mux[b]);
        }
    }
    MTM_UNROLL
    for (int a = 0; a < DM::LM; ++a) {
        if (a < l) {
            MTM_UNROLL
            for (int b = a; b < DM::LM; ++b)
                if (b < l) Sx[a][b] = p.Lam(a * l + b);
        }
    }
    MTM_UNROLL
    for (int j = 0; j < DM::LM; ++j) {
        if (j < l) {
            const double y = xt[j];
            const bool obs = finite_bits(y) && ((finmask >> j) & 1u);
            const double s = Sx[j][j];
            const double r = obs ? fast_rcp(s) : 0.0;
            const double v = obs ? (y - mux[j]) : 0.0;
            f.q = fma(v * v, r, f.q);
            if (obs) f.mant *= s;  // predicated DMUL (a select against 1.0 costs four IMAD moves, which steal FP64 issue slots)
            f.k += obs ? 1 : 0;
            MTM_UNROLL
            for (int a = j + 1; a < DM::LM; ++a) {
                if (a < l) {
                    const double g = Sx[j][a] * r;
                    mux[a] = fma(g, v, mux[a]);
                    MTM_UNROLL
                    for (int b = a; b < DM::LM; ++b)
                        if (b < l) Sx[a][b] = fma(-g, Sx[j][b], Sx[a][b]);
                }
            }
        }
    }
}

// ---- rank-one shortcuts --------------------------------------------------------------------------------------
// When at most ONE hidden-state coordinate of a step is missing (coordinate j, different from lane to lane),
// conditioning on the present ones leaves P = sigma2 e_j e_j' with exact zeros everywhere else (and mu_i = z_i for
// i != j).  The generic code would then push zeros through C = P H, the gains, A'PA; these two functions do only
// the arithmetic on the one live direction, in the generic code's operation order, so every number has the same
// bits: row j of H / A is picked with the 0/1 indicator as a multiplicand (exact: one term, the others are zeros),
// C[j][b] = sigma2 H[j][b], Sx = Lambda + h_j c', P' = Gamma + a_j (sigma2 a_j)'.  `miss` = bit mask of the
// missing coordinate (0 or one bit set per lane).  Taken on a warp-uniform vote like the point-mass shortcut; at
// 20 % element-wise missingness 41 % of the steps have exactly one missing coordinate (41 % none).

template <class DM>
__device__ __forceinline__ double miss_ind(uint32_t miss, int i) {
    // 1.0 if coordinate i is the missing one else 0.0 (built on the integer pipe)
    return __hiloint2double(((miss >> i) & 1u) ? 0x3ff00000 : 0, 0);
}

// (mu, P) <- (mu A, Gamma + A' P A) for P = sigma2 e_j e_j'
template <class DM, class PV>
__device__ __forceinline__ void filter_propagate_r1(const DM &dm, FilterState<DM> &f, const PV &p, uint32_t miss) {
    const int d = dm.d();
    double nmu[DM::DM], aj[DM::DM];
    MTM_UNROLL
    for (int j = 0; j < DM::DM; ++j) nmu[j] = aj[j] = 0.0;
    double s2 = 0.0;
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            const double mi = f.mu[i];
            const double ei = miss_ind<DM>(miss, i);
            s2 += f.P[i][i];
            MTM_UNROLL
            for (int j = 0; j < DM::DM; ++j) {
                if (j < d) {
                    nmu[j] = fma(mi, p.A(i * d + j), nmu[j]);
                    aj[j] = fma(ei, p.A(i * d + j), aj[j]);
                }
            }
        }
    }
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            f.mu[i] = nmu[i];
            const double ai = aj[i];
            MTM_UNROLL
            for (int j = i; j < DM::DM; ++j)
                if (j < d) f.P[i][j] = fma(ai, s2 * aj[j], p.G(i * d + j));
        }
    }
}

// predictive of x_t and conditioning on its present coordinates for P = sigma2 e_j e_j'
template <class DM, class PV>
__device__ __forceinline__ void filter_condition_x_r1(const DM &dm, FilterState<DM> &f, const PV &p,
                                                      const double (&xt)[DM::LM], uint32_t finmask, uint32_t miss,
                                                      const bool LAST) {
    const int d = dm.d(), l = dm.l();
    double mux[DM::LM], c[DM::LM], hj[DM::LM];
    double Sx[DM::LM][DM::LM];
    MTM_UNROLL
    for (int b = 0; b < DM::LM; ++b) mux[b] = c[b] = hj[b] = 0.0;
    double muj = 0.0, s2 = 0.0;  // mean and variance of the one uncertain coordinate
    MTM_UNROLL
    for (int i = 0; i < DM::DM; ++i) {
        if (i < d) {
            const double mi = f.mu[i], pii = f.P[i][i];
            const double ei = miss_ind<DM>(miss, i);
            muj = fma(ei, mi, muj);
            s2 += pii;
            MTM_UNROLL
            for (int b = 0; b < DM::LM; ++b) {
                if (b < l) {
                    mux[b] = fma(mi, p.H(i * l + b), mux[b]);
                    c[b] = fma(pii, p.H(i * l + b), c[b]);
                    hj[b] = fma(ei, p.H(i * l + b), hj[b]);
                }
            }
        }
    }
    MTM_UNROLL
    for (int a = 0; a < DM::LM; ++a) {
        if (a < l) {
            MTM_UNROLL
            for (int b = a; b < DM::LM; ++b)
                if (b < l) Sx[a][b] = fma(hj[a], c[b], p.Lam(a * l + b));
        }
    }
    MTM_UNROLL
    for (int j = 0; j < DM::LM; ++j) {
        if (j < l) {
            const double y = xt[j];
            const bool obs = finite_bits(y) && ((finmask >> j) & 1u);
            const double s = Sx[j][j];
            const double r = obs ? fast_rcp(s) : 0.0;
            const double v = obs ? (y - mux[j]) : 0.0;
            f.q = fma(v * v, r, f.q);
            if (obs) f.mant *= s;  // predicated DMUL (see filter_condition_z)
            f.k += obs ? 1 : 0;
            const double g = c[j] * r;
            if (!LAST) {
                muj = fma(g, v, muj);
                s2 = fma(-g, c[j], s2);
            }
            MTM_UNROLL
            for (int b = j + 1; b < DM::LM; ++b)
                if (b < l) c[b] = fma(-g, Sx[j][b], c[b]);
            MTM_UNROLL
            for (int a = j + 1; a < DM::LM; ++a) {
                if (a < l) {
                    const double gx = Sx[j][a] * r;
                    mux[a] = fma(gx, v, mux[a]);
                    MTM_UNROLL
                    for (int b = a; b < DM::LM; ++b)
                        if (b < l) Sx[a][b] = fma(-gx, Sx[j][b], Sx[a][b]);
                }
            }
        }
    }
    if (!LAST) {
        MTM_UNROLL
        for (int i = 0; i < DM::DM; ++i) {
            if (i < d) {
                const bool mine = (miss >> i) & 1u;
                f.mu[i] = mine ? muj : f.mu[i];
                f.P[i][i] = mine ? s2 : f.P[i][i];  // the present coordinates keep their exact zero
            }
        }
    }
}

// keep the running determinant in range: move the binary exponent of `mant` into `expo`
template <class DM>
__device__ __forceinline__ void filter_rescale(FilterState<DM> &f) {
    const int hi = __double2hiint(f.mant);
    const int e = (hi >> 20) & 0x7ff;
    if (e != 0 && e != 0x7ff) {  // normal number (zero / subnormal / inf / NaN are left alone)
        f.expo += e - 1022;
        f.mant = __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(f.mant));
    }
}

// -1/2 (k log 2 pi + log det) - 1/2 q.  A negative determinant gives NaN like the reference's
// log((2 pi)^k det) (:771-772); no present coordinate gives -0.0.
//
// Failure flags are derived here instead of per pivot: a NaN/inf pivot makes the determinant or the
// quadratic form non-finite (NONFINITE: the reference's det/solve raise on such a block); an exactly
// zero pivot makes the determinant exactly zero (SINGULAR).  A finite negative determinant is NOT an
// error (the reference returns NaN for it without raising).
template <class DM>
__device__ __forceinline__ double filter_finish(const FilterState<DM> &f, int &bad) {
    bad |= (!finite_bits(f.mant) || !finite_bits(f.q)) ? MTM_STATUS_NONFINITE : 0;
    bad |= (f.mant == 0.0) ? MTM_STATUS_SINGULAR : 0;
    const double logdet = fma((double)f.expo, 0.6931471805599453094, log(f.mant));
    return -0.5 * fma((double)f.k, 1.8378770664093454836, logdet) - 0.5 * f.q;
}

// vector load of N doubles (16-byte loads when the row size allows; pointers are 16-byte aligned
// by the ABI contract when N is even)
template <int NM>
__device__ __forceinline__ void load_row(const double *__restrict__ p, double (&v)[NM], int n_rt, bool is_static) {
    if (is_static && (NM % 2 == 0)) {
        MTM_UNROLL
        for (int k = 0; k < NM / 2; ++k) {
            const double2 t = __ldg(reinterpret_cast<const double2 *>(p) + k);
            v[2 * k] = t.x;
            v[2 * k + 1] = t.y;
        }
    } else {
        MTM_UNROLL
        for (int k = 0; k < NM; ++k)
            if (k < n_rt) v[k] = __ldg(p + k);
    }
}

// One full evaluation.  z_i / x_i point at step 0 of trajectory i; zs / xs = step strides in doubles.
// fin[t] = bit mask of finite unconditional means at step t: bits [0, d) for z, [16, 16 + l) for x.
// `tshift` is a kernel argument that is ZERO at run time.  Its only use: in the observation-only
// instantiation (HAS_Z = false) the hidden-state code stays in the kernel behind `tshift != 0`, a
// warp-uniform branch that is never taken.  With the same code shape as the joint kernel ptxas 12.9 keeps the
// uniform-datapath parameter operands there too (modelled FP64-pipe efficiency 0.89 instead of 0.80,
// tools/sass_reuse.py); the null z pointer is never dereferenced.
// PREFIX = true (estep_prefix.cu only): after every step t the log-likelihood of the first t+1 steps is
// written to pre[t * pre_stride] — the sequential form has every prefix likelihood for free
// (cluster_propensities_over_time, MMM:532-577, needs all of them).
template <class DM, bool HAS_Z, class PS, bool PREFIX = false>
__device__ __forceinline__ double filter_eval(const DM &dm, const double *__restrict__ z_i, int64_t zs,
                                              const double *__restrict__ x_i, int64_t xs, int T0,
                                              const PS &psrc, const uint32_t *fin, int &bad, int tshift,
                                              double *pre = nullptr, int64_t pre_stride = 0) {
    FilterState<DM> f;
    {
        const ParamView<DM, PS> p0(dm, psrc);
        filter_init(dm, f, p0);
    }
    // Loads: x_t is requested at the top of its own step (first used after the propagation and the
    // z-conditioning); z_{t+1} is prefetched at the top of step t into zn and handed over at the loop end.
    // That hand-over is PINNED to the loop end by a predicate ptxas cannot resolve (always true: the exponent
    // counter never reaches INT_MAX).  As plain copies ptxas 12.9 hoisted them to right behind the loads to
    // recycle the registers, so every step stalled for a full memory latency on the first copy (ncu: 7.7 % of
    // the warp samples of the time loop on one IMAD.MOV, long scoreboard).  Loading z where it is used, or
    // anywhere but the loop top, makes ptxas drop the uniform-datapath parameter operands (LDCU -> LDC,
    // tools/sass_loops.py), a far bigger loss.
    // The observation-only instantiation has no z stream, so there the same pattern is applied to x.
    constexpr bool XPRE = !HAS_Z;
    double zt[DM::DM], xt[DM::LM];
    if (HAS_Z || tshift != 0) load_row<DM::DM>(z_i, zt, dm.d(), DM::kStatic);
    if (XPRE) load_row<DM::LM>(x_i, xt, dm.l(), DM::kStatic);
    const ParamView<DM, PS> p(dm, psrc);
    bool p_point = false;  // warp-uniform: the posterior of the previous step is a point mass (P == 0, mu == z)
    bool p_r1 = false;     // warp-uniform: ... has at most one uncertain coordinate per lane (rank-one shortcuts)
    uint32_t p_miss = 0;   // per lane: bit mask of the coordinates of the previous z_t that were missing
    for (int t = 0; t < T0; ++t) {
        const bool last = (t + 1 == T0);  // warp-uniform
        const int tn = last ? t : t + 1;
        double zn[DM::DM], xn[DM::LM];
        if (HAS_Z || tshift != 0) load_row<DM::DM>(z_i + (int64_t)tn * zs, zn, dm.d(), DM::kStatic);
        if (XPRE) load_row<DM::LM>(x_i + (int64_t)tn * xs, xn, dm.l(), DM::kStatic);
        else load_row<DM::LM>(x_i + (int64_t)t * xs, xt, dm.l(), DM::kStatic);
        const uint32_t fm = fin[t];
        // warp-uniform votes: are all hidden-state coordinates of this step present for all 32 trajectories
        // (point mass), or all but at most one per trajectory (rank one)?
        bool z_point = false, z_r1 = false;
        uint32_t zmiss = 0;
        if (HAS_Z || tshift != 0) {
            MTM_UNROLL
            for (int j = 0; j < DM::DM; ++j)
                if (j < dm.d()) zmiss |= ((finite_bits(zt[j]) ? 1u : 0u) & (fm >> j) & 1u) ? 0u : (1u << j);
z_point = __all_sync(0xffffffffu, zmiss == 0u);
            // (both votes unconditional: a short-circuited second vote made ptxas 12.9 drop the uniform-datapath
            // parameter operands of the whole time loop; the rank-one code exists in the joint kernel only — the
            // observation-only instantiation keeps its z code as the never-taken shape donor it was before)
            if (HAS_Z) z_r1 = __all_sync(0xffffffffu, (zmiss & (zmiss - 1u)) == 0u);  // includes the point case
        }
        if (t > 0) {
            if (p_point) filter_propagate_point(dm, f, p);
            else if (HAS_Z && p_r1) filter_propagate_r1(dm, f, p, p_miss);
            else filter_propagate(dm, f, p);
        }
        if (HAS_Z || tshift != 0) filter_condition_z(dm, f, zt, fm & 0xffffu);
        if (z_point) filter_condition_x_point(dm, f, p, xt, fm >> 16);
        else if (HAS_Z && z_r1) filter_condition_x_r1(dm, f, p, xt, fm >> 16, zmiss, last);
        else filter_condition_x(dm, f, p, xt, fm >> 16, last);
        p_r1 = z_r1;
        p_miss = zmiss;
        p_point = z_point;
        filter_rescale(f);
        if (PREFIX) pre[(int64_t)t * pre_stride] = filter_finish(f, bad);
#define MTM_PINNED_COPY(dst, src)                                                                   \
    asm volatile("{ .reg .pred pp; setp.ne.s32 pp, %2, 0x7fffffff; @pp mov.f64 %0, %1; }" \
                 : "+d"(dst) : "d"(src), "r"(f.expo))
        if (HAS_Z) {
            MTM_UNROLL
            for (int k = 0; k < DM::DM; ++k)
                if (k < dm.d()) MTM_PINNED_COPY(zt[k], zn[k]);
        } else {
            MTM_UNROLL
            for (int k = 0; k < DM::DM; ++k) zt[k] = zn[k];
            MTM_UNROLL
            for (int k = 0; k < DM::LM; ++k)
                if (k < dm.l()) MTM_PINNED_COPY(xt[k], xn[k]);
        }
#undef MTM_PINNED_COPY
    }
    return filter_finish(f, bad);
}

// finiteness masks of the unconditional stacked mean [m A^t | m A^t H] for one cluster; run by one
// thread per cluster in a kernel prologue.  T0 words.
template <class DM, class PS>
__device__ void mean_finite_masks(const DM &dm, const PS &psrc, int T0, uint32_t *fin) {
    const ParamView<DM, PS> p(dm, psrc);
    const int d = dm.d(), l = dm.l();
    double mu[MTM_MAX_DIM], nm[MTM_MAX_DIM];
    for (int i = 0; i < d; ++i) mu[i] = p.m(i);
    for (int t = 0; t < T0; ++t) {
        if (t > 0) {
            for (int j = 0; j < d; ++j) {
                double s = 0.0;
                for (int i = 0; i < d; ++i) s = fma(mu[i], p.A(i * d + j), s);
                nm[j] = s;
            }
            for (int j = 0; j < d; ++j) mu[j] = nm[j];
        }
        uint32_t m = 0;
        for (int j = 0; j < d; ++j) m |= finite_bits(mu[j]) ? (1u << j) : 0u;
        for (int b = 0; b < l; ++b) {
            double s = 0.0;
            for (int i = 0; i < d; ++i) s = fma(mu[i], p.H(i * l + b), s);
            m |= finite_bits(s) ? (1u << (16 + b)) : 0u;
        }
        fin[t] = m;
    }
}

// The same masks, eight lanes per cluster (lane j = state coordinate j and observation coordinate j): the sums run in
// the same order as above, so the masks are identical.  The lane's columns of A and H are read ONCE, from the
// parameter rows in GLOBAL memory: lanes of one warp want different parameters, and a constant-bank load with 32
// different addresses is replayed 32 times — the one-thread-per-cluster prologue and a first lane-parallel version
// that still read the constant bank both cost ~10 us per block (ncu: 0.7 % of a launch with 15 tiles per block,
// 5 % with the 2 tiles per block of a 1.25 M shard on 8 GPUs).  Every lane of the warp must call it; `c_ok` = this
// 8-lane group's cluster exists (the others compute on a clamped cluster and write nothing).
template <class DM>
__device__ __forceinline__ void mean_finite_masks_lanes(const DM &dm, const double *__restrict__ row, int T0,
                                                        uint32_t *fin, bool c_ok) {
    const ParamView<DM, SmemSrc> p(dm, SmemSrc{row});
    const int d = dm.d(), l = dm.l();
    const int lane = threadIdx.x & 31, j = lane & 7, base = lane & 24;
    double a_col[MTM_MAX_DIM], h_col[MTM_MAX_DIM];
    MTM_UNROLL
    for (int i = 0; i < MTM_MAX_DIM; ++i) {
        a_col[i] = (i < d && j < d) ? __ldg(row + p.oA + i * d + j) : 0.0;
        h_col[i] = (i < d && j < l) ? __ldg(row + p.oH + i * l + j) : 0.0;
    }
    double mu = j < d ? __ldg(row + 1 + j) : 0.0;
    for (int t = 0; t < T0; ++t) {
        if (t > 0) {
            double s = 0.0;
            MTM_UNROLL
            for (int i = 0; i < MTM_MAX_DIM; ++i) {
                const double mi = __shfl_sync(0xffffffffu, mu, base + i);
                if (i < d) s = fma(mi, a_col[i], s);
            }
            mu = s;
        }
        double sx = 0.0;
        MTM_UNROLL
        for (int i = 0; i < MTM_MAX_DIM; ++i) {
            const double mi = __shfl_sync(0xffffffffu, mu, base + i);
            if (i < d) sx = fma(mi, h_col[i], sx);
        }
        const uint32_t zb = (__ballot_sync(0xffffffffu, j < d && finite_bits(mu)) >> base) & 0xffu;
        const uint32_t xb = (__ballot_sync(0xffffffffu, j < l && finite_bits(sx)) >> base) & 0xffu;
        if (j == 0 && c_ok) fin[t] = zb | (xb << 16);
    }
}

}  // namespace mtm
