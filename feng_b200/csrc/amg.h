// Aggregation multigrid hierarchy (amg.cu) and the preconditioners built on it (precond.cu).  Internal header.
#pragma once
#include <cstdint>
#include <vector>

#include "system.h"

namespace b200 {

constexpr int     AMG_MAX_LEVELS  = 12;
constexpr int     AMG_MAX_DENSE   = 384; // coarsest level is inverted densely below this size
constexpr int     AMG_CHEB_DEGREE = 2;
constexpr double  AMG_CHEB_RATIO  = 4.;  // smoother targets [lambda_max / ratio, lambda_max] of D^-1 A
constexpr uint8_t AMG_FLD_P       = 100; // field id of the pressure rows (velocity components are 0 .. dim-1)
constexpr uint8_t AMG_FLD_NONE    = 255; // row outside every field / ghost row of another rank

struct AmgLevel {
  int64_t        n = 0, nnz = 0;
  const int64_t *ia = nullptr;
  const int32_t *ja = nullptr;
  const double  *val = nullptr;
  int64_t       *ia_own = nullptr; // coarse levels own their CSR arrays; level 0 aliases the system matrix
  int32_t       *ja_own = nullptr;
  double        *val_own = nullptr;
  bool           is_system = false, decoupled = false;
  bool           halo = false; // several GPUs: products on this level update the ghost entries of their input first
  // level 0 only: single-precision copy of the entries the hierarchy works on (active rows, columns of the same field):
  // the smoother, the residual and the Galerkin product of level 0 stream 8 bytes per entry of a third of the matrix
  // instead of 12 bytes of all of it
  int64_t *ia_s = nullptr;
  int32_t *ja_s = nullptr;
  float   *val_s = nullptr;
  int64_t  nnz_s = 0;
  double        *dinv = nullptr;
  double         lam = 1.;
  bool           have_eig = false;
  // transfer to the next level: up to two parents per unknown (kind 1: weight 1; kind 2: weight 1/2 each; 0: inactive)
  int32_t *par0 = nullptr, *par1 = nullptr;
  uint8_t *pkind = nullptr;
  // work vectors
  double *x = nullptr, *b = nullptr, *t = nullptr, *d = nullptr, *ev = nullptr;
};

struct Amg {
  std::vector<AmgLevel> L;
  const uint8_t        *d_fld = nullptr;     // level-0 field id per row (owned by the preconditioner)
  uint8_t              *d_active0 = nullptr; // level-0 row mask when level 0 is aggregated directly (P1 spaces)
  int                   dense_n = 0;
  double               *dense = nullptr, *cinv = nullptr, *d_nrm = nullptr;
  bool                  symbolic = false, verbose = false;
  int                   cheb_degree = AMG_CHEB_DEGREE, cycles = 1; // B200_AMG_DEGREE / B200_AMG_CYCLES
  int                   pre_degree = 0;                            // B200_AMG_PRE: degree of the pre-smoother (0 = cheb_degree)
  double                cheb_ratio = AMG_CHEB_RATIO;               // B200_AMG_RATIO
  int                   gamma_from = 1;                            // first level whose coarse level is visited gamma times
  int                   mis_distance = 1;                          // B200_AMG_MIS: roots of the aggregates = maximal independent set of distance 2 (or 1)
  int                   gamma = 2;                                 // B200_AMG_GAMMA: cycle index below the finest level (2 = W-cycle)
  double                overcorrect = 1.5;                          // B200_AMG_OVERCORRECT: scaling of the aggregation-level corrections
  bool                  use_f32 = true;                            // B200_AMG_F32=0: level 0 works on the FP64 system matrix
  // several GPUs: the hierarchy is rank-local down to level `gl` - 1 (block-Jacobi across ranks: couplings to ghost columns are
  // dropped there); from level `gl` (= 2: the first aggregation level of a P2 hierarchy) on it is GLOBAL.  The level-gl operator
  // of all ranks, INCLUDING the couplings across the partition cuts (Galerkin contributions of the rows that read ghost
  // columns), is gathered as (row, column, value) triplets, merged into one CSR matrix that every rank holds, and coarsened /
  // smoothed / solved redundantly by a second hierarchy `G`; a cycle all-reduces the level-gl right-hand side, runs G and keeps
  // its own slice.  Smooth error spanning several sub-domains is removed at the resolution of level gl (6-8 fine cells), which is
  // what keeps the iteration count of a chain of sub-domains from growing with its length.
  bool     gc_active = false;
  int      gl = 0;                            // the global level
  int64_t  gc_n = 0, gc_off = 0, gc_nloc = 0; // global / first own / own number of level-gl unknowns
  int32_t *gc_p0 = nullptr, *gc_p1 = nullptr; // [n fine] global level-gl ids of the (up to two) vertex parents, -1 = none
  uint8_t *gc_kind = nullptr;                 // [n fine] 0 none, 1 vertex unknown, 2 mid-edge unknown (weights 1 / one half)
  const uint8_t *d_fld_all = nullptr;         // field map with the ghost rows still labelled (owned by the preconditioner)
  Amg     *G = nullptr;                       // replicated hierarchy on the merged level-gl matrix
  const uint8_t *d_fcol0 = nullptr, *d_kcol0 = nullptr; // column labels used by the level-0 single-precision copy
  int64_t *g_ia = nullptr;
  int32_t *g_ja = nullptr;
  double  *g_val = nullptr;
  int64_t  g_nnz = 0;
  double  *gc_b = nullptr, *gc_x = nullptr;   // [gc_n]
  uint64_t *t_key = nullptr, *t_keyall = nullptr; // triplets: own (capacity t_cap) / gathered (world * t_cap)
  double   *t_val = nullptr, *t_valall = nullptr;
  int64_t   t_cap = 0;
};

void amg_free(Amg *A);
int  amg_build_field_map(System *S, uint8_t **d_fld);
int  amg_mask_ghosts(System *S, uint8_t *d_fld);
int  amg_setup_symbolic(System *S, Amg *A, const uint8_t *d_fld, int fld_lo, int fld_hi, int space, const uint8_t *d_fld_all = nullptr);
int  amg_setup_numeric(System *S, Amg *A);
int  amg_vcycle(System *S, Amg *A, const double *b, double *x);

// precond.cu: B200_PC_AMG / B200_PC_SCHUR_AMG
int  precond_setup(System *S, int pc);
int  precond_apply(System *S, int pc, const double *r, double *z);
void precond_free(System *S);
// degrade the smoother (Chebyshev -> damped Jacobi) after GMRES stagnation; false if there is nothing left to degrade
bool precond_fallback(System *S);

} // namespace b200
