// =============================================================================
// TEST INFRASTRUCTURE ONLY.  Harness around the UNMODIFIED reference (arthurbawin/feNG,
// compiled by oracle/Makefile from the sources where they lie under /root/reference).
// It exposes, through a small C ABI usable from ctypes, everything the parity tests need
// from the reference's own CPU hot path:
//   * DOF addressing tables, CSR pattern (feEZCompressedRowStorage), colours, basis and
//     quadrature tables, Jacobians,
//   * per-element matrices / residuals (feBilinearForm::computeMatrix/computeResidual),
//   * the assembled CSR values + rhs through a restatement of the Pardiso-backend scatter
//     (src/feLinearSystemMklPardiso.cpp:501-749; that class is compiled only with MKL),
//   * essential-component constraints (src/feLinearSystemMklPardiso.cpp:994-1117),
//   * a full Newton solve through the reference's own solveNewtonRaphson with a stub
//     feLinearSystem backend (Eigen SparseLU) -- used for solution parity and as the timed
//     CPU baseline ("reference" kind) in bench.py.
// Nothing in feng_b200/ (the product) links or loads this file.
// =============================================================================
#include "feAPI.h"
#ifdef WITH_B200_ADAPTER
#include "../adapter/feLinearSystemB200.h" // the product's C++ drop-in backend, driven by the unmodified Newton loop
#endif

#include <Eigen/Sparse>
#include <Eigen/SparseLU>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#if defined(HAVE_OMP)
#include <omp.h>
#endif

extern std::vector<double> solAtTimeN; // src/feNonLinearSolver.cpp:35 (read by feBilinearForm::initialize)
extern int                 FE_VERBOSE;

// ---------------------------------------------------------------------------------------
// Recipes
// ---------------------------------------------------------------------------------------
extern "C" {
typedef struct {
  int    kind;        // 0 scalar diffusion+source, 1 Stokes div-form, 2 NS div-form, 3 NS Laplacian-form,
                      // 4 Stokes Laplacian-form, 5 CHNS, 6 / 7 Stokes Poiseuille channel, divergence / Laplacian form
                      // (tests/withLinearSolver/stokes.cpp:183-275), 8 scalar diffusion periodic between the channel's
                      // inlet and outlet (feSpace::setPeriodic*, src/GenericSolver.cpp:187-194), 9 scalar diffusion with a
                      // space-dependent diffusivity callback
  int    order;       // polynomial order of the primary field (pressure uses order-1)
  int    quad_degree; // quadrature degree given to createFiniteElementSpace
  int    field;       // analytic field family: 0 polynomial MMS of the reference tests, 1 Kovasznay(Re=1/mu),
                      // 2 zero fields / unit source
  double mu;          // viscosity or diffusivity
  double rho;         // density factor on the convective and transient terms
  int    transient;   // 1: add the transient (vector) mass form, c0 taken from ref_set_solution
  int    p_essential; // 1: pressure essential on the whole boundary "Bord" (reference MMS tests),
                      // 0: pressure essential on "PointPression" if that entity exists, else free
} ref_recipe_t;
}

namespace {

struct Params {
  int    field;
  double mu, rho;
  int    kind;
};

const double PI = 3.14159265358979323846;

// ---- analytic fields ---------------------------------------------------------------------------
// field 0: the polynomial manufactured solution of tests/withLinearSolver/navier_stokes.cpp:19-52 and
//          (scalar) tests/withLinearSolver/convergenceLaplace.cpp:18-32, generalised to (mu, rho).
// field 1: Kovasznay flow, Re = rho/mu.
void uSolCb(const feFunctionArguments &args, const std::vector<double> &par, std::vector<double> &res)
{
  const double x = args.pos[0], y = args.pos[1];
  const int field = (int)par[0];
  if(field == 0) {
    res[0] = pow(x, 4) * pow(y, 4);
    res[1] = -4. / 5. * pow(x, 3) * pow(y, 5);
  } else if(field == 1) {
    const double Re = par[2] / par[1];
    const double lam = Re / 2. - sqrt(Re * Re / 4. + 4. * PI * PI);
    res[0] = 1. - exp(lam * x) * cos(2. * PI * y);
    res[1] = lam / (2. * PI) * exp(lam * x) * sin(2. * PI * y);
  } else {
    res[0] = 0.;
    res[1] = 0.;
  }
}

double pSolCb(const feFunctionArguments &args, const std::vector<double> &par)
{
  const double x = args.pos[0], y = args.pos[1];
  const int field = (int)par[0];
  if(field == 0) return x * x * y * y;
  if(field == 1) {
    const double Re = par[2] / par[1];
    const double lam = Re / 2. - sqrt(Re * Re / 4. + 4. * PI * PI);
    return 0.5 * par[2] * (1. - exp(2. * lam * x));
  }
  return 0.;
}

// Source of the momentum equation written as in the reference tests (everything multiplied by -1):
//   -rho (u.grad)u - grad p + mu lap(u) + f = 0   ->   uSrc = -( -rho u.grad u - grad p + mu lap u )
// par = {field, mu, rho, withConvection}
void uSrcCb(const feFunctionArguments &args, const std::vector<double> &par, std::vector<double> &res)
{
  const double x = args.pos[0], y = args.pos[1];
  const int    field = (int)par[0];
  const double mu = par[1], rho = par[2], conv = par[3];
  if(field == 0) {
    const double minus_dpdx[2] = {-2. * x * y * y, -2. * x * x * y};
    const double lap_u[2]      = {12. * (x * x * y * y * y * y + x * x * x * x * y * y),
                                  -4. / 5. * (6. * x * y * y * y * y * y + 20. * x * x * x * y * y * y)};
    const double u[2]          = {x * x * x * x * y * y * y * y, (-4. / 5. * x * x * x * y * y * y * y * y)};
    const double gradu[2][2]   = {{4. * x * x * x * y * y * y * y, -12. * x * x * y * y * y * y * y / 5.},
                                  {4. * x * x * x * x * y * y * y, -4. * x * x * x * y * y * y * y}};
    const double uDotGradu[2]  = {u[0] * gradu[0][0] + u[1] * gradu[1][0], u[0] * gradu[0][1] + u[1] * gradu[1][1]};
    res[0] = -(-conv * rho * uDotGradu[0] + minus_dpdx[0] + mu * lap_u[0]);
    res[1] = -(-conv * rho * uDotGradu[1] + minus_dpdx[1] + mu * lap_u[1]);
  } else {
    res[0] = 0.;
    res[1] = 0.;
  }
}

// Poiseuille channel flow, tests/withLinearSolver/stokes.cpp:165-181: par = {H or L, dpdx}
void poiseuilleUCb(const feFunctionArguments &args, const std::vector<double> &par, std::vector<double> &res)
{
  const double y = args.pos[1];
  res[0] = -par[1] / 2. * y * (par[0] - y);
  res[1] = 0.;
}
double poiseuillePCb(const feFunctionArguments &args, const std::vector<double> &par)
{
  return -par[1] * (par[0] - args.pos[0]);
}
// periodic-in-x scalar field on the channel [0,5] x [0,1] and its source for -k lap(u) (kind 8)
double periodicSolCb(const feFunctionArguments &args, const std::vector<double> &)
{
  const double x = args.pos[0], y = args.pos[1];
  return sin(2. * PI * x / 5.) * y * (1. - y) + y * y;
}
double periodicSrcCb(const feFunctionArguments &args, const std::vector<double> &par)
{
  const double x = args.pos[0], y = args.pos[1], w = 2. * PI / 5.;
  // -(u_xx + u_yy) k, sign convention of the reference's Source form (convergenceLaplace.cpp:25-32: source = +k lap u ... )
  const double lap = -w * w * sin(w * x) * y * (1. - y) - 2. * sin(w * x) + 2.;
  return par[0] * lap;
}
// space-dependent diffusivity of kind 9 (field-dependent coefficient forms, tests/withLinearSolver/scalarFE.cpp)
double varDiffusivityCb(const feFunctionArguments &args, const std::vector<double> &par)
{
  const double x = args.pos[0], y = args.pos[1];
  return par[0] * (1. + 0.5 * x + 0.25 * y * y);
}

double sSolCb(const feFunctionArguments &args, const std::vector<double> &par)
{
  const double x = args.pos[0], y = args.pos[1], z = args.pos[2];
  const int field = (int)par[0];
  if(field == 0) return pow(x, 6) + pow(y, 6) + pow(z, 6);
  if(field == 4) return sin(PI * x) * sin(PI * y); // tests/withLinearSolver/scalarFE.cpp:14-21
  return 0.;
}

double sSrcCb(const feFunctionArguments &args, const std::vector<double> &par)
{
  const double x = args.pos[0], y = args.pos[1], z = args.pos[2];
  const int    field = (int)par[0];
  const double k = par[1];
  if(field == 0) return k * 30. * (pow(x, 4) + pow(y, 4) + pow(z, 4));
  if(field == 4) return -2. * k * PI * PI * sin(PI * x) * sin(PI * y); // tests/withLinearSolver/scalarFE.cpp:23-28
  return -1.;
}

// Expose the protected pattern arrays of the reference's own pattern builder.
struct PatternPeek : public feEZCompressedRowStorage {
  using feEZCompressedRowStorage::feEZCompressedRowStorage;
  const std::vector<feInt> &ia() const { return ia_Pardiso; }
  const std::vector<feInt> &ja() const { return ja_Pardiso; }
};

// ---------------------------------------------------------------------------------------
// Stub CPU backend: restates feLinearSystemMklPardiso's assembly/constraint semantics, solves with
// the vendored Eigen SparseLU.  Driven by the reference's unmodified Newton loop.
// ---------------------------------------------------------------------------------------
class feLinearSystemStub : public feLinearSystem
{
public:
  feInt               _nInc;
  feInt               _nnz;
  std::vector<feInt>  _ia, _ja;
  std::vector<double> _val, _rhs, _du;
  std::vector<feInt>  _rowsToConstrain;
  bool                _constraintInit = false;
  double              _tAsmMat = 0., _tAsmRes = 0., _tSolve = 0.;
  int                 _nAsmMat = 0, _nAsmRes = 0, _nSolve = 0;

  feLinearSystemStub(const std::vector<feBilinearForm *> forms, const feMetaNumber *numbering)
    : feLinearSystem(forms, numbering)
  {
    _recomputeMatrix = true;
    _nInc            = numbering->getNbUnknowns();
    PatternPeek crs((int)_nInc, _formMatrices, _numMatrixForms, numbering);
    _ia  = crs.ia();
    _ja  = crs.ja();
    _nnz = crs.getNumNNZ();
    _val.assign(_nnz, 0.);
    _rhs.assign(_nInc, 0.);
    _du.assign(_nInc, 0.);
  }
  ~feLinearSystemStub() {}

  feInt getSystemSize() const { return _nInc; }
  void  getRHSMaxNorm(double *norm) const
  {
    double m = 0.;
    for(double v : _rhs) m = fmax(m, fabs(v));
    *norm = m;
  }
  void getResidualMaxNorm(double *norm) const
  {
    double m = 0.;
    for(double v : _du) m = fmax(m, fabs(v));
    *norm = m;
  }
  void setToZero()
  {
    if(_recomputeMatrix) setMatrixToZero();
    setResidualToZero();
  }
  void setMatrixToZero() { std::fill(_val.begin(), _val.end(), 0.); }
  void setResidualToZero() { std::fill(_rhs.begin(), _rhs.end(), 0.); }
  void assemble(const feSolution *sol, const bool onlyTransient = false)
  {
    if(_recomputeMatrix) assembleMatrices(sol, onlyTransient);
    assembleResiduals(sol);
  }

  // Restatement of src/feLinearSystemMklPardiso.cpp:524-663 (colour loop, essential filter, sorted scatter).
  void assembleMatrices(const feSolution *sol, const bool onlyTransient = false)
  {
    auto t0 = std::chrono::steady_clock::now();
    for(feInt eq = 0; eq < _numMatrixForms; ++eq) {
      feBilinearForm *f0 = _formMatrices[eq];
      if(onlyTransient && !f0->isTransientMatrix()) continue;
      const feCncGeo *cnc       = f0->getCncGeo();
      const int       numColors = cnc->getNbColor();
      for(int iColor = 0; iColor < numColors; ++iColor) {
        const std::vector<int> &listElmC = cnc->getListElmPerColorI(iColor);
        const int               nE       = (int)listElmC.size();
#if defined(HAVE_OMP)
#pragma omp parallel for schedule(dynamic)
#endif
        for(int iElm = 0; iElm < nE; ++iElm) {
#if defined(HAVE_OMP)
          feBilinearForm *f = _formMatrices[eq + omp_get_thread_num() * _numMatrixForms];
#else
          feBilinearForm *f = f0;
#endif
          f->computeMatrix(sol, listElmC[iElm]);
          const double *const      *Ae   = f->getAe();
          const std::vector<feInt> &adrI = f->getAdrI();
          const std::vector<feInt> &adrJ = f->getAdrJ();
          for(size_t i = 0; i < adrI.size(); ++i) {
            const feInt I = adrI[i];
            if(I >= _nInc) continue;
            const feInt  beg = _ia[I], end = _ia[I + 1];
            const feInt *row = _ja.data() + beg;
            for(size_t j = 0; j < adrJ.size(); ++j) {
              const feInt J = adrJ[j];
              if(J >= _nInc) continue;
              const feInt *pos = std::lower_bound(row, row + (end - beg), J);
              _val[beg + (pos - row)] += Ae[i][j];
            }
          }
        }
      }
    }
    _tAsmMat += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    _nAsmMat++;
  }

  // Restatement of src/feLinearSystemMklPardiso.cpp:699-741.
  void assembleResiduals(const feSolution *sol)
  {
    auto t0 = std::chrono::steady_clock::now();
    for(feInt eq = 0; eq < _numResidualForms; ++eq) {
      feBilinearForm *f0        = _formResiduals[eq];
      const feCncGeo *cnc       = f0->getCncGeo();
      const int       numColors = cnc->getNbColor();
      for(int iColor = 0; iColor < numColors; ++iColor) {
        const std::vector<int> &listElmC = cnc->getListElmPerColorI(iColor);
        const int               nE       = (int)listElmC.size();
#if defined(HAVE_OMP)
#pragma omp parallel for schedule(dynamic)
#endif
        for(int iElm = 0; iElm < nE; ++iElm) {
#if defined(HAVE_OMP)
          feBilinearForm *f = _formResiduals[eq + omp_get_thread_num() * _numResidualForms];
#else
          feBilinearForm *f = f0;
#endif
          f->computeResidual(sol, listElmC[iElm]);
          const double             *Be   = f->getBe();
          const std::vector<feInt> &adrI = f->getAdrI();
          for(size_t i = 0; i < adrI.size(); ++i)
            if(adrI[i] < _nInc) _rhs[adrI[i]] += Be[i];
        }
      }
    }
    _tAsmRes += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    _nAsmRes++;
  }

  // Row list as in src/feLinearSystemMklPardiso.cpp:1005-1041.
  void initConstraintRows(const feSolution *sol)
  {
    _constraintInit = true;
    std::vector<feInt> adr;
    for(const auto &space : sol->_spaces) {
      const int nComponents = space->getNumComponents();
      const int nFunctions  = space->getNumFunctions();
      const int nElm        = space->getNumElements();
      if(nComponents > 1) {
        adr.resize(nFunctions, 0);
        for(int i = 0; i < nComponents; ++i) {
          if(space->isEssentialComponent(i)) {
            for(int iElm = 0; iElm < nElm; ++iElm) {
              space->initializeAddressingVector(iElm, adr);
              for(int j = 0; j < nFunctions; ++j)
                if(j % nComponents == i && adr[j] < _nInc) _rowsToConstrain.push_back(adr[j]);
            }
          }
        }
      }
    }
    std::sort(_rowsToConstrain.begin(), _rowsToConstrain.end());
  }

  // Semantics of src/feLinearSystemMklPardiso.cpp:1092-1114: zero column, zero row, unit diagonal, zero rhs.
  void constrainEssentialComponents(const feSolution *sol)
  {
    if(!_constraintInit) initConstraintRows(sol);
    if(_rowsToConstrain.empty()) return;
    std::vector<char> flag(_nInc, 0);
    for(feInt r : _rowsToConstrain) flag[r] = 1;
    for(feInt i = 0; i < _nInc; ++i)
      for(feInt k = _ia[i]; k < _ia[i + 1]; ++k) {
        if(flag[_ja[k]]) _val[k] = 0.;
        if(flag[i]) _val[k] = (_ja[k] == i) ? 1. : 0.;
      }
    for(feInt r : _rowsToConstrain) _rhs[r] = 0.;
  }

  // src/feLinearSystemMklPardiso.cpp:1119-1149
  void applyPeriodicity()
  {
    for(const auto &pair : _numbering->PeriodicDOF()) {
      const int masterDOF = pair.first, slaveDOF = pair.second;
      if(slaveDOF < _nInc && masterDOF < _nInc) {
        for(feInt k = _ia[slaveDOF]; k < _ia[slaveDOF + 1]; ++k) {
          _val[k] = 0.;
          if(_ja[k] == slaveDOF) _val[k] = 1.;
          if(_ja[k] == masterDOF) _val[k] = -1.;
        }
        _rhs[slaveDOF] = 0.;
      }
    }
  }
  void permute() {}

  bool solve(double *normDx, double *normResidual, double *normAxb, int *nIter)
  {
    auto t0 = std::chrono::steady_clock::now();
    typedef Eigen::SparseMatrix<double, Eigen::RowMajor, long> SpMat;
    Eigen::Map<const SpMat> A(_nInc, _nInc, _nnz, _ia.data(), _ja.data(), _val.data());
    Eigen::SparseMatrix<double, Eigen::ColMajor, long> Ac = A;
    Eigen::SparseLU<Eigen::SparseMatrix<double, Eigen::ColMajor, long>> lu;
    lu.compute(Ac);
    if(lu.info() != Eigen::Success) return false;
    Eigen::Map<Eigen::VectorXd> b(_rhs.data(), _nInc);
    Eigen::VectorXd             x = lu.solve(b);
    if(lu.info() != Eigen::Success) return false;
    Eigen::VectorXd r = A * x - b;
    for(feInt i = 0; i < _nInc; ++i) _du[i] = x[i];
    *normDx       = x.cwiseAbs().maxCoeff();
    *normResidual = b.cwiseAbs().maxCoeff();
    *normAxb      = r.cwiseAbs().maxCoeff();
    *nIter        = 0;
    _tSolve += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    _nSolve++;
    return true;
  }

  // src/feLinearSystemMklPardiso.cpp:761-793
  void correctSolution(feSolution *sol, const bool correctSolutionDot = false)
  {
    std::vector<double> &v = correctSolutionDot ? sol->getSolutionDot() : sol->getSolution();
    for(feInt i = 0; i < _nInc; ++i) v[i] += _du[i];
  }
  void assignResidualToDCResidual(feSolutionContainer *c)
  {
    for(feInt i = 0; i < _nInc; ++i) c->_fResidual[0][i] = _rhs[i];
  }
  void applyCorrectionToResidual(double coeff, std::vector<double> &d)
  {
    for(feInt i = 0; i < _nInc; ++i) _rhs[i] += coeff * d[i];
  }
  void viewMatrix() const {}
  void viewRHS() const {}
  void viewResidual() const {}
  void writeMatrix(const std::string, const double) {}
  void writeRHS(const std::string, const double) {}
  void writeResidual(const std::string, const double) {}
};

// ---- CHNS (config 5): parameters and property callbacks ------------------------------------------------
// g_chns = {rhoA, rhoB, viscA, viscB, mobility, sigma, epsilon, fx, fy, Su0, Su1, Sp, Sphi, Smu, limiter, degenerateMobility,
//           phiOrder, formulation (0 CHNS_Abels, 1 CHNS_MassAveraged, 2 CHNS_Khanwale), alpha, Re, Pe, Cn, We, Fr, rhoA, rhoB};
//           set by ref_set_chns_params before ref_create(kind = 5).  The property laws are the ones
// CHNS_Solver hands to its weak form (src/CHNS_Solver.cpp:124-235): linear mixing in phi, optional clipping of phi to
// [-1, 1], constant or degenerate mobility M |1 - phi^2|.
double g_chns[26] = {1., 1., 1., 1., 1., 1., 0.1, 0., 0., 0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 1., 1., 1., 1., 1., 1., 1.};

double chnsLinearCb(const feFunctionArguments &args, const std::vector<double> &par)
{
  const double phi = par[2] != 0. ? fmax(-1., fmin(1., args.u)) : args.u;
  return (par[0] - par[1]) / 2. * phi + (par[0] + par[1]) / 2.;
}
double chnsSlopeCb(const feFunctionArguments &, const std::vector<double> &par) { return (par[0] - par[1]) / 2.; }
double chnsMobilityCb(const feFunctionArguments &args, const std::vector<double> &par)
{
  return par[1] != 0. ? par[0] * fabs(1. - args.u * args.u) : par[0];
}
void chnsVecConstCb(const feFunctionArguments &, const std::vector<double> &par, std::vector<double> &res)
{
  res[0] = par[0];
  res[1] = par[1];
}
// initial phase marker / potential (smooth, |phi| crosses 1 so that the limiter is exercised)
double chnsPhiCb(const feFunctionArguments &args, const std::vector<double> &)
{
  const double x = args.pos[0], y = args.pos[1];
  return 1.2 * cos(PI * x) * cos(PI * y);
}
double chnsMuCb(const feFunctionArguments &args, const std::vector<double> &)
{
  const double x = args.pos[0], y = args.pos[1];
  return 0.3 * sin(PI * x) * sin(2. * PI * y) + 0.1 * x;
}


struct RefProblem {
  ref_recipe_t rc;
  feMesh2DP1  *mesh = nullptr;
  int          dim  = 2;

  std::vector<feFunction *>       sfun;
  std::vector<feVectorFunction *> vfun;

  std::vector<feSpace *>        spaces, essentialSpaces, interior; // interior = spaces used by the forms
  std::vector<feSpace *>        extra;                             // created but possibly not part of `spaces`
  feMetaNumber                 *numbering = nullptr;
  feSolution                   *sol       = nullptr;
  std::vector<feBilinearForm *> forms;
  feLinearSystemStub           *sys = nullptr;
  feSpace                      *uSpace = nullptr, *pSpace = nullptr;
  feVectorFunction             *uExact = nullptr;
  feFunction                   *pExact = nullptr, *sExact = nullptr;
  // state at the previous time step handed to feBilinearForm::initialize through the global solAtTimeN
  // (ref_set_solution_n); empty = the current solution
  std::vector<double>           solN;
  const std::vector<double>    &stateN() const { return solN.empty() ? sol->getSolution() : solN; }

  ~RefProblem()
  {
    delete sys;
    for(auto *f : forms) delete f;
    delete sol;
    delete numbering;
    for(auto *s : extra)
      if(std::find(spaces.begin(), spaces.end(), s) == spaces.end()) delete s;
    for(auto *s : spaces) delete s;
    for(auto *f : sfun) delete f;
    for(auto *f : vfun) delete f;
    delete mesh;
  }
};

feFunction *mkS(RefProblem *P, ScalarField f, std::vector<double> par)
{
  P->sfun.push_back(new feFunction(f, par));
  return P->sfun.back();
}
feVectorFunction *mkV(RefProblem *P, VectorField f, std::vector<double> par)
{
  P->vfun.push_back(new feVectorFunction(f, par));
  return P->vfun.back();
}

bool hasEntity(feMesh *mesh, const std::string &name)
{
  for(auto *cnc : mesh->getCncGeo())
    if(cnc->getID() == name) return true;
  return false;
}

#define CHK(x)                                                                                                         \
  do {                                                                                                                 \
    if((x) != FE_STATUS_OK) {                                                                                          \
      delete P;                                                                                                        \
      return nullptr;                                                                                                  \
    }                                                                                                                  \
  } while(0)

} // namespace

extern "C" {
int ref_error_norms(void *h, const double *sol, double *out);

// CHNS parameter block for the next ref_create(kind = 5); see g_chns
void ref_set_chns_params(const double *p, int n)
{
  g_chns[17] = g_chns[18] = 0.;
  for(int i = 19; i < 26; ++i) g_chns[i] = 1.;
  for(int i = 0; i < n && i < 26; ++i) g_chns[i] = p[i];
}

int ref_max_threads()
{
#if defined(HAVE_OMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void ref_set_threads(int n)
{
#if defined(HAVE_OMP)
  if(n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void *ref_create(const char *meshFile, const ref_recipe_t *rc)
{
  setVerbose(0);
  RefProblem *P = new RefProblem;
  P->rc         = *rc;
  P->mesh       = new feMesh2DP1(std::string(meshFile));
  P->dim        = P->mesh->getDim();
  const double fld = (double)rc->field;

  if(rc->kind == 0) {
    // Scalar diffusion + source: tests/withLinearSolver/convergenceLaplace.cpp:64-77, exe/example1.cpp:133-170
    feFunction *sol = mkS(P, sSolCb, {fld});
    feFunction *src = mkS(P, sSrcCb, {fld, rc->mu});
    feFunction *k   = mkS(P, constantCallback, {rc->mu});
    P->sExact       = sol;
    feSpace *uB = nullptr, *u = nullptr;
    CHK(createFiniteElementSpace(uB, P->mesh, elementType::LAGRANGE, rc->order, "U", "Bord", rc->quad_degree, sol));
    CHK(createFiniteElementSpace(u, P->mesh, elementType::LAGRANGE, rc->order, "U", "Domaine", rc->quad_degree,
                                 &scalarConstant::zero));
    P->spaces          = {u, uB};
    P->essentialSpaces = {uB};
    P->interior        = {u};
    P->uSpace          = u;
    P->numbering       = new feMetaNumber(P->mesh, P->spaces, P->essentialSpaces);
    P->sol             = new feSolution(P->numbering->getNbDOFs(), P->spaces, P->essentialSpaces);
    feBilinearForm *diff = nullptr, *source = nullptr;
    if(P->dim == 2)
      CHK(createBilinearForm(diff, {u}, new feSysElm_Diffusion<2>(k)));
    else
      CHK(createBilinearForm(diff, {u}, new feSysElm_Diffusion<3>(k)));
    CHK(createBilinearForm(source, {u}, new feSysElm_Source(src)));
    P->forms = {diff, source};
    if(rc->transient) {
      feBilinearForm *mass = nullptr;
      feFunction     *r    = mkS(P, constantCallback, {rc->rho});
      CHK(createBilinearForm(mass, {u}, new feSysElm_TransientMass(r)));
      P->forms.push_back(mass);
    }
  } else if(rc->kind == 5) {
    // Cahn-Hilliard Navier-Stokes, Abels et al. (CHNS_Solver, src/CHNS_Solver.cpp:236-420): fields U (P2 vector), P (P1),
    // Phi, Mu (P1 or P2), ONE monolithic weak form whose Jacobian is computed by finite differences
    const double     *g    = g_chns;
    feVectorFunction *uSol = mkV(P, uSolCb, {fld, rc->mu, rc->rho});
    feFunction       *pSol = mkS(P, pSolCb, {fld, rc->mu, rc->rho});
    feFunction       *fSol = mkS(P, chnsPhiCb, {});
    feFunction       *mSol = mkS(P, chnsMuCb, {});
    P->uExact              = uSol;
    P->pExact              = pSol;
    const int fo = (int)g[16];
    feSpace  *u = nullptr, *uB = nullptr, *p = nullptr, *pB = nullptr, *phi = nullptr, *mu = nullptr;
    CHK(createFiniteElementSpace(u, P->mesh, elementType::VECTOR_LAGRANGE, 2, "U", "Domaine", rc->quad_degree, uSol));
    CHK(createFiniteElementSpace(uB, P->mesh, elementType::VECTOR_LAGRANGE, 2, "U", "Bord", rc->quad_degree, uSol));
    CHK(createFiniteElementSpace(p, P->mesh, elementType::LAGRANGE, 1, "P", "Domaine", rc->quad_degree, pSol));
    CHK(createFiniteElementSpace(phi, P->mesh, elementType::LAGRANGE, fo, "Phi", "Domaine", rc->quad_degree, fSol));
    CHK(createFiniteElementSpace(mu, P->mesh, elementType::LAGRANGE, fo, "Mu", "Domaine", rc->quad_degree, mSol));
    P->spaces          = {u, uB, p, phi, mu};
    P->essentialSpaces = {uB};
    if(hasEntity(P->mesh, "PointPression")) {
      CHK(createFiniteElementSpace(pB, P->mesh, elementType::LAGRANGE, 0, "P", "PointPression", rc->quad_degree, pSol));
      P->spaces.push_back(pB);
      P->essentialSpaces.push_back(pB);
    }
    P->interior  = {u, p, phi, mu};
    P->uSpace    = u;
    P->pSpace    = p;
    P->numbering = new feMetaNumber(P->mesh, P->spaces, P->essentialSpaces);
    P->sol       = new feSolution(P->numbering->getNbDOFs(), P->spaces, P->essentialSpaces);
    feFunction       *rho   = mkS(P, chnsLinearCb, {g[0], g[1], g[14]});
    feFunction       *drho  = mkS(P, chnsSlopeCb, {g[0], g[1]});
    feFunction       *visc  = mkS(P, chnsLinearCb, {g[2], g[3], g[14]});
    feFunction       *dvisc = mkS(P, chnsSlopeCb, {g[2], g[3]});
    feFunction       *mob   = mkS(P, chnsMobilityCb, {g[4], g[15]});
    feVectorFunction *force = mkV(P, chnsVecConstCb, {g[7], g[8]});
    feVectorFunction *srcU  = mkV(P, chnsVecConstCb, {g[9], g[10]});
    feFunction       *srcP  = mkS(P, constantCallback, {g[11]});
    feFunction       *srcF  = mkS(P, constantCallback, {g[12]});
    feFunction       *srcM  = mkS(P, constantCallback, {g[13]});
    feBilinearForm     *chns = nullptr;
    if(g[17] == 1.) {
      // src/CHNS_Solver.cpp:398-416: {mass_alpha, surfaceTension, epsilon}
      std::vector<double> prm = {g[18], g[5], g[6]};
      CHK(createBilinearForm(chns, {u, p, phi, mu},
                             new CHNS_MassAveraged<2>(rho, drho, visc, dvisc, mob, force, srcP, srcU, srcF, srcM, prm)));
    } else if(g[17] == 2.) {
      // src/CHNS_Solver.cpp:418-446: {Re, Pe, Cn, We, Fr, rhoA, rhoB}
      std::vector<double> prm = {g[19], g[20], g[21], g[22], g[23], g[24], g[25]};
      CHK(createBilinearForm(chns, {u, p, phi, mu},
                             new CHNS_Khanwale<2>(rho, drho, visc, dvisc, mob, force, srcP, srcU, srcF, srcM, prm)));
    } else {
      std::vector<double> prm = {g[5], g[6]};
      CHK(createBilinearForm(chns, {u, p, phi, mu},
                             new CHNS_Abels<2>(rho, drho, visc, dvisc, mob, force, srcP, srcU, srcF, srcM, prm)));
    }
    P->forms.push_back(chns);
  } else if(rc->kind == 6 || rc->kind == 7) {
    // Stokes Poiseuille, tests/withLinearSolver/stokes.cpp:183-275; data/poiseuille0.msh names its entities
    // Domain/Inlet/Outlet/NoSlip, data/poiseuille1.msh Domaine/Entree/Sortie/NoSlip
    const bool        divForm = rc->kind == 6;
    const std::string dom = hasEntity(P->mesh, "Domain") ? "Domain" : "Domaine";
    const std::string in  = hasEntity(P->mesh, "Inlet") ? "Inlet" : "Entree";
    const std::string out = hasEntity(P->mesh, "Outlet") ? "Outlet" : "Sortie";
    const double      H = 1., L = 5., dpdx = -1.0;
    feVectorFunction *uSol = mkV(P, poiseuilleUCb, {H, dpdx});
    feFunction       *pSol = mkS(P, poiseuillePCb, {L, dpdx});
    P->uExact              = uSol;
    P->pExact              = pSol;
    feSpace *u = nullptr, *p = nullptr, *uInlet = nullptr, *uOutlet = nullptr, *uNoSlip = nullptr;
    CHK(createFiniteElementSpace(u, P->mesh, elementType::VECTOR_LAGRANGE, rc->order, "U", dom, rc->quad_degree, &vectorConstant::zero));
    CHK(createFiniteElementSpace(uInlet, P->mesh, elementType::VECTOR_LAGRANGE, rc->order, "U", in, rc->quad_degree, uSol));
    CHK(createFiniteElementSpace(uOutlet, P->mesh, elementType::VECTOR_LAGRANGE, rc->order, "U", out, rc->quad_degree, &vectorConstant::zero));
    CHK(createFiniteElementSpace(uNoSlip, P->mesh, elementType::VECTOR_LAGRANGE, rc->order, "U", "NoSlip", rc->quad_degree, &vectorConstant::zero));
    CHK(createFiniteElementSpace(p, P->mesh, elementType::LAGRANGE, rc->order - 1, "P", dom, rc->quad_degree, &scalarConstant::zero));
    P->spaces          = {uInlet, uNoSlip, u, p};
    P->essentialSpaces = {uNoSlip, uInlet};
    if(divForm) {
      P->spaces.push_back(uOutlet);
      uOutlet->setEssentialComponent(1, true);
    }
    P->extra     = {uOutlet};
    P->interior  = {u, p};
    P->uSpace    = u;
    P->pSpace    = p;
    P->numbering = new feMetaNumber(P->mesh, P->spaces, P->essentialSpaces);
    P->sol       = new feSolution(P->numbering->getNbDOFs(), P->spaces, P->essentialSpaces);
    feBilinearForm *divSigma = nullptr, *diffU = nullptr, *gradP = nullptr, *divU = nullptr;
    CHK(createBilinearForm(divU, {p, u}, new feSysElm_MixedDivergence<2>(&scalarConstant::one)));
    P->forms = {divU};
    if(divForm) {
      CHK(createBilinearForm(divSigma, {u, p}, new feSysElm_DivergenceNewtonianStress<2>(&scalarConstant::one, &scalarConstant::one)));
      P->forms.push_back(divSigma);
    } else {
      CHK(createBilinearForm(gradP, {u, p}, new feSysElm_MixedGradient<2>(&scalarConstant::minusOne)));
      CHK(createBilinearForm(diffU, {u}, new feSysElm_VectorDiffusion<2>(&scalarConstant::minusOne, &scalarConstant::one)));
      P->forms.push_back(diffU);
      P->forms.push_back(gradP);
    }
  } else if(rc->kind == 8 || rc->kind == 9) {
    // kind 8: scalar diffusion on the channel mesh, essential on the walls, PERIODIC between inlet (master) and outlet
    //         (slave, offset (5, 0, 0)): exercises the pattern extras of feEZCompressedRowStorage
    //         (src/feCompressedRowStorage.cpp:96-107) and applyPeriodicity (src/feLinearSystemMklPardiso.cpp:1119-1149)
    // kind 9: scalar diffusion with a space-dependent diffusivity callback on any mesh with Domaine / Bord
    const bool        periodic = rc->kind == 8;
    const std::string dom = hasEntity(P->mesh, "Domain") ? "Domain" : "Domaine";
    feFunction *sol = periodic ? mkS(P, periodicSolCb, {}) : mkS(P, sSolCb, {fld});
    feFunction *src = periodic ? mkS(P, periodicSrcCb, {rc->mu}) : mkS(P, sSrcCb, {fld, rc->mu});
    feFunction *k   = periodic ? mkS(P, constantCallback, {rc->mu}) : mkS(P, varDiffusivityCb, {rc->mu});
    P->sExact       = sol;
    feSpace *u = nullptr;
    CHK(createFiniteElementSpace(u, P->mesh, elementType::LAGRANGE, rc->order, "U", dom, rc->quad_degree, &scalarConstant::zero));
    if(periodic) {
      const std::string in  = hasEntity(P->mesh, "Inlet") ? "Inlet" : "Entree";
      const std::string out = hasEntity(P->mesh, "Outlet") ? "Outlet" : "Sortie";
      feSpace *uW = nullptr, *uIn = nullptr, *uOut = nullptr;
      CHK(createFiniteElementSpace(uW, P->mesh, elementType::LAGRANGE, rc->order, "U", "NoSlip", rc->quad_degree, sol));
      CHK(createFiniteElementSpace(uIn, P->mesh, elementType::LAGRANGE, rc->order, "U", in, rc->quad_degree, &scalarConstant::zero));
      CHK(createFiniteElementSpace(uOut, P->mesh, elementType::LAGRANGE, rc->order, "U", out, rc->quad_degree, &scalarConstant::zero));
      uIn->setPeriodic(true);
      uOut->setPeriodic(true);
      uIn->setPeriodicMaster(true);
      uOut->setPeriodicSlave(true);
      uIn->setMatchingPeriodicSpace(uOut);
      uOut->setMatchingPeriodicSpace(uIn);
      uIn->setPeriodicOffset({5., 0., 0.});
      uOut->setPeriodicOffset({5., 0., 0.});
      P->spaces          = {u, uW, uIn, uOut};
      P->essentialSpaces = {uW};
    } else {
      feSpace *uB = nullptr;
      CHK(createFiniteElementSpace(uB, P->mesh, elementType::LAGRANGE, rc->order, "U", "Bord", rc->quad_degree, sol));
      P->spaces          = {u, uB};
      P->essentialSpaces = {uB};
    }
    P->interior  = {u};
    P->uSpace    = u;
    P->numbering = new feMetaNumber(P->mesh, P->spaces, P->essentialSpaces);
    P->sol       = new feSolution(P->numbering->getNbDOFs(), P->spaces, P->essentialSpaces);
    feBilinearForm *diff = nullptr, *source = nullptr;
    CHK(createBilinearForm(diff, {u}, new feSysElm_Diffusion<2>(k)));
    CHK(createBilinearForm(source, {u}, new feSysElm_Source(src)));
    P->forms = {diff, source};
  } else {
    // (Navier-)Stokes Taylor-Hood: tests/withLinearSolver/navier_stokes.cpp:63-99, stokes.cpp
    const bool withConv = (rc->kind == 2 || rc->kind == 3);
    const bool divForm  = (rc->kind == 1 || rc->kind == 2);
    feVectorFunction *uSol = mkV(P, uSolCb, {fld, rc->mu, rc->rho});
    feFunction       *pSol = mkS(P, pSolCb, {fld, rc->mu, rc->rho});
    feVectorFunction *uSrc = mkV(P, uSrcCb, {fld, rc->mu, rc->rho, withConv ? 1. : 0.});
    feFunction       *mu   = mkS(P, constantCallback, {rc->mu});
    feFunction       *mRho = mkS(P, constantCallback, {-rc->rho});
    feFunction       *rho  = mkS(P, constantCallback, {rc->rho});
    P->uExact              = uSol;
    P->pExact              = pSol;

    feSpace *u = nullptr, *uB = nullptr, *p = nullptr, *pB = nullptr;
    CHK(createFiniteElementSpace(u, P->mesh, elementType::VECTOR_LAGRANGE, rc->order, "U", "Domaine", rc->quad_degree,
                                 uSol));
    CHK(createFiniteElementSpace(uB, P->mesh, elementType::VECTOR_LAGRANGE, rc->order, "U", "Bord", rc->quad_degree,
                                 uSol));
    CHK(createFiniteElementSpace(p, P->mesh, elementType::LAGRANGE, rc->order - 1, "P", "Domaine", rc->quad_degree,
                                 pSol));
    P->spaces          = {u, uB, p};
    P->essentialSpaces = {uB};
    if(rc->p_essential) {
      CHK(createFiniteElementSpace(pB, P->mesh, elementType::LAGRANGE, rc->order - 1, "P", "Bord", rc->quad_degree,
                                   pSol));
      P->spaces.push_back(pB);
      P->essentialSpaces.push_back(pB);
    } else if(hasEntity(P->mesh, "PointPression")) {
      CHK(createFiniteElementSpace(pB, P->mesh, elementType::LAGRANGE, 0, "P", "PointPression", rc->quad_degree,
                                   pSol));
      P->spaces.push_back(pB);
      P->essentialSpaces.push_back(pB);
    }
    P->interior  = {u, p};
    P->uSpace    = u;
    P->pSpace    = p;
    P->numbering = new feMetaNumber(P->mesh, P->spaces, P->essentialSpaces);
    P->sol       = new feSolution(P->numbering->getNbDOFs(), P->spaces, P->essentialSpaces);

    feBilinearForm *convU = nullptr, *divSigma = nullptr, *diffU = nullptr, *gradP = nullptr, *divU = nullptr,
                   *source = nullptr, *mass = nullptr;
    if(withConv) {
      CHK(createBilinearForm(convU, {u}, new feSysElm_VectorConvectiveAcceleration<2>(mRho)));
      P->forms.push_back(convU);
    }
    CHK(createBilinearForm(divU, {p, u}, new feSysElm_MixedDivergence<2>(&scalarConstant::one)));
    CHK(createBilinearForm(source, {u}, new feSysElm_VectorSource<2>(uSrc)));
    P->forms.push_back(divU);
    P->forms.push_back(source);
    if(divForm) {
      CHK(createBilinearForm(divSigma, {u, p}, new feSysElm_DivergenceNewtonianStress<2>(&scalarConstant::one, mu)));
      P->forms.push_back(divSigma);
    } else {
      CHK(createBilinearForm(gradP, {u, p}, new feSysElm_MixedGradient<2>(&scalarConstant::minusOne)));
      CHK(createBilinearForm(diffU, {u}, new feSysElm_VectorDiffusion<2>(&scalarConstant::minusOne, mu)));
      P->forms.push_back(diffU);
      P->forms.push_back(gradP);
    }
    if(rc->transient) {
      CHK(createBilinearForm(mass, {u}, new feSysElm_TransientVectorMass<2>(mRho)));
      P->forms.push_back(mass);
    }
  }

  P->sol->initialize(P->mesh);
  solAtTimeN = P->stateN();
  P->sys     = new feLinearSystemStub(P->forms, P->numbering);
  return P;
}

void ref_destroy(void *h) { delete(RefProblem *)h; }

// out[0..15]: dim, nVertices, nElm, nVertPerElm, nDOF, nInc, nnz, nQuad, nColors, nInteriorSpaces, nForms,
//             numMatrixForms
int ref_info(void *h, int64_t *out)
{
  RefProblem     *P   = (RefProblem *)h;
  const feCncGeo *cnc = P->uSpace->getCncGeo();
  out[0]              = P->dim;
  out[1]              = (int64_t)P->mesh->getVertices().size();
  out[2]              = cnc->getNumElements();
  out[3]              = cnc->getNumVerticesPerElem();
  out[4]              = P->numbering->getNbDOFs();
  out[5]              = P->numbering->getNbUnknowns();
  out[6]              = P->sys->_nnz;
  out[7]              = P->uSpace->getNumQuadPoints();
  out[8]              = cnc->getNbColor();
  out[9]              = (int64_t)P->interior.size();
  out[10]             = (int64_t)P->forms.size();
  int nm              = 0;
  for(auto *f : P->forms) nm += f->hasMatrix() ? 1 : 0;
  out[11] = nm;
  return 0;
}

int ref_get_mesh(void *h, double *xyz, int32_t *conn)
{
  RefProblem *P = (RefProblem *)h;
  auto       &V = P->mesh->getVertices();
  for(size_t i = 0; i < V.size(); ++i) {
    xyz[3 * i + 0] = V[i].x();
    xyz[3 * i + 1] = V[i].y();
    xyz[3 * i + 2] = V[i].z();
  }
  const feCncGeo *cnc = P->uSpace->getCncGeo();
  const int       nv  = cnc->getNumVerticesPerElem();
  for(int e = 0; e < cnc->getNumElements(); ++e)
    for(int j = 0; j < nv; ++j) conn[nv * e + j] = cnc->getVertexConnectivity(e, j);
  return 0;
}

// out: nFunctions, nComponents
int ref_space_info(void *h, int s, int64_t *out)
{
  RefProblem *P = (RefProblem *)h;
  out[0]        = P->interior[s]->getNumFunctions();
  out[1]        = P->interior[s]->getNumComponents();
  return 0;
}

int ref_get_adr(void *h, int s, int64_t *adr)
{
  RefProblem        *P  = (RefProblem *)h;
  feSpace           *S  = P->interior[s];
  const int          nF = S->getNumFunctions();
  std::vector<feInt> a(nF);
  for(int e = 0; e < S->getNumElements(); ++e) {
    S->initializeAddressingVector(e, a);
    for(int j = 0; j < nF; ++j) adr[(int64_t)nF * e + j] = a[j];
  }
  return 0;
}

// Scalar tables [k][i] for scalar spaces; for vector spaces the [k][i][c] layout of the reference is returned
// (nQuad * nF * nC doubles).  dLdt only meaningful in 3D.
int ref_get_tables(void *h, int s, double *L, double *dLdr, double *dLds, double *dLdt)
{
  RefProblem *P  = (RefProblem *)h;
  feSpace    *S  = P->interior[s];
  const int   nF = S->getNumFunctions(), nC = S->getNumComponents(), nQ = S->getNumQuadPoints();
  const int   n  = nF * nC;
  if(nC == 1) {
    for(int k = 0; k < nQ; ++k)
      for(int i = 0; i < nF; ++i) {
        L[n * k + i]    = S->getFunctionAtQuadNode(i, k);
        dLdr[n * k + i] = S->getdFunctiondrAtQuadNode(i, k);
        dLds[n * k + i] = P->dim >= 2 ? S->getdFunctiondsAtQuadNode(i, k) : 0.;
        dLdt[n * k + i] = P->dim >= 3 ? S->getdFunctiondtAtQuadNode(i, k) : 0.;
      }
  } else {
    // Evaluate from shape functions at the quadrature nodes (vector-valued L: nF x nC per point)
    const std::vector<double> &r = S->getRQuadraturePoints(), &ss = S->getSQuadraturePoints(),
                              &t = S->getTQuadraturePoints();
    for(int k = 0; k < nQ; ++k) {
      double              rr[3] = {r[k], ss[k], t[k]};
      std::vector<double> l = S->L(rr), a = S->dLdr(rr), b = S->dLds(rr);
      for(int i = 0; i < n; ++i) {
        L[n * k + i]    = l[i];
        dLdr[n * k + i] = a[i];
        dLds[n * k + i] = b[i];
        dLdt[n * k + i] = 0.;
      }
    }
  }
  return 0;
}

int ref_get_quadrature(void *h, double *w, double *r, double *s, double *t)
{
  RefProblem *P  = (RefProblem *)h;
  feSpace    *S  = P->uSpace;
  const int   nQ = S->getNumQuadPoints();
  for(int k = 0; k < nQ; ++k) {
    w[k] = S->getQuadratureWeights()[k];
    r[k] = S->getRQuadraturePoints()[k];
    s[k] = S->getSQuadraturePoints()[k];
    t[k] = S->getTQuadraturePoints()[k];
  }
  return 0;
}

int ref_get_jacobians(void *h, double *J)
{
  RefProblem                *P = (RefProblem *)h;
  const std::vector<double> &j = P->uSpace->getCncGeo()->getJacobians();
  std::memcpy(J, j.data(), j.size() * sizeof(double));
  return 0;
}

int ref_get_colors(void *h, int32_t *elmToColor)
{
  RefProblem             *P = (RefProblem *)h;
  const std::vector<int> &c = P->uSpace->getCncGeo()->getColorElm();
  for(size_t i = 0; i < c.size(); ++i) elmToColor[i] = c[i];
  return 0;
}

int ref_get_pattern(void *h, int64_t *ia, int64_t *ja)
{
  RefProblem *P = (RefProblem *)h;
  for(size_t i = 0; i < P->sys->_ia.size(); ++i) ia[i] = P->sys->_ia[i];
  for(size_t i = 0; i < P->sys->_ja.size(); ++i) ja[i] = P->sys->_ja[i];
  return 0;
}

int ref_get_solution(void *h, double *sol, double *solDot)
{
  RefProblem *P = (RefProblem *)h;
  const int   n = P->sol->getNumDOFs();
  std::memcpy(sol, P->sol->getSolution().data(), n * sizeof(double));
  std::memcpy(solDot, P->sol->getSolutionDot().data(), n * sizeof(double));
  return 0;
}

int ref_set_solution(void *h, const double *sol, const double *solDot, double c0, double t)
{
  RefProblem *P = (RefProblem *)h;
  const int   n = P->sol->getNumDOFs();
  std::memcpy(P->sol->getSolution().data(), sol, n * sizeof(double));
  if(solDot) std::memcpy(P->sol->getSolutionDot().data(), solDot, n * sizeof(double));
  P->sol->setC0(c0);
  P->sol->setCurrentTime(t);
  solAtTimeN = P->stateN();
  return 0;
}

// state at the previous time step (NULL: back to "equal to the current solution") and the time step
int ref_set_solution_n(void *h, const double *solN, double dt)
{
  RefProblem *P = (RefProblem *)h;
  {
    // feSolution::_dt has no setter of its own: initializeTemporalSolution (src/feSolution.cpp:97-104) sets it
    const double t = P->sol->getCurrentTime();
    if(dt > 0.) P->sol->initializeTemporalSolution(t, t + dt, 1);
    P->sol->setCurrentTime(t);
  }
  if(solN)
    P->solN.assign(solN, solN + P->sol->getNumDOFs());
  else
    P->solN.clear();
  return 0;
}

// out: M, N, hasMatrix, elementSystemType id, isTransientMatrix
int ref_form_info(void *h, int f, int64_t *out)
{
  RefProblem *P = (RefProblem *)h;
  out[0]        = P->forms[f]->getLocalMatrixM();
  out[1]        = P->forms[f]->getLocalMatrixN();
  out[2]        = P->forms[f]->hasMatrix() ? 1 : 0;
  out[3]        = (int64_t)P->forms[f]->getID();
  out[4]        = P->forms[f]->isTransientMatrix() ? 1 : 0;
  return 0;
}

// Element matrix (M x N row-major, zeros if the form has no matrix) and residual of ONE form on ONE element,
// straight from feBilinearForm::computeMatrix / computeResidual.
int ref_element(void *h, int f, int elem, double *Ae, double *Be, int64_t *adrI, int64_t *adrJ)
{
  RefProblem     *P = (RefProblem *)h;
  feBilinearForm *F = P->forms[f];
  const feInt     M = F->getLocalMatrixM(), N = F->getLocalMatrixN();
  solAtTimeN        = P->stateN();
  if(F->hasMatrix()) {
    F->computeMatrix(P->sol, elem);
    const double *const *A = F->getAe();
    for(feInt i = 0; i < M; ++i)
      for(feInt j = 0; j < N; ++j) Ae[i * N + j] = A[i][j];
  } else {
    for(feInt i = 0; i < M * N; ++i) Ae[i] = 0.;
  }
  F->computeResidual(P->sol, elem);
  for(feInt i = 0; i < M; ++i) Be[i] = F->getBe()[i];
  for(feInt i = 0; i < M; ++i) adrI[i] = F->getAdrI()[i];
  for(feInt j = 0; j < N; ++j) adrJ[j] = F->getAdrJ()[j];
  return 0;
}

// what: bit 0 residual, bit 1 matrix.  Zeroes first.  seconds[0] = matrix time, seconds[1] = residual time.
int ref_assemble(void *h, int what, double *values, double *rhs, double *seconds)
{
  RefProblem *P = (RefProblem *)h;
  solAtTimeN    = P->stateN();
  double t0m = P->sys->_tAsmMat, t0r = P->sys->_tAsmRes;
  if(what & 2) {
    P->sys->setMatrixToZero();
    P->sys->assembleMatrices(P->sol, false);
    if(values) std::memcpy(values, P->sys->_val.data(), P->sys->_nnz * sizeof(double));
  }
  if(what & 1) {
    P->sys->setResidualToZero();
    P->sys->assembleResiduals(P->sol);
    if(rhs) std::memcpy(rhs, P->sys->_rhs.data(), P->sys->_nInc * sizeof(double));
  }
  if(seconds) {
    seconds[0] = P->sys->_tAsmMat - t0m;
    seconds[1] = P->sys->_tAsmRes - t0r;
  }
  return 0;
}

// Apply constrainEssentialComponents + applyPeriodicity to the currently assembled system and return it.
int ref_constrain(void *h, double *values, double *rhs)
{
  RefProblem *P = (RefProblem *)h;
  P->sys->constrainEssentialComponents(P->sol);
  P->sys->applyPeriodicity();
  std::memcpy(values, P->sys->_val.data(), P->sys->_nnz * sizeof(double));
  std::memcpy(rhs, P->sys->_rhs.data(), P->sys->_nInc * sizeof(double));
  return 0;
}

// Direct solve of the currently assembled (and constrained) system; du has nInc entries.
int ref_solve_current(void *h, double *du, double *norms)
{
  RefProblem *P = (RefProblem *)h;
  int         it;
  bool        ok = P->sys->solve(&norms[0], &norms[1], &norms[2], &it);
  std::memcpy(du, P->sys->_du.data(), P->sys->_nInc * sizeof(double));
  return ok ? 0 : -1;
}

// Full stationary solve through the reference's unmodified createTimeIntegrator -> solveNewtonRaphson
// (tests/withLinearSolver/navier_stokes.cpp:110-121).  out: errU (or scalar error), errP, tAsmMat, tAsmRes, tSolve,
// nAsmMat, nAsmRes, nSolve.  sol_out receives the nDOF solution.
int ref_newton(void *h, double tolRes, double tolCor, int maxIter, double *sol_out, double *out)
{
  RefProblem       *P = (RefProblem *)h;
  feNLSolverOptions NL{tolRes, tolCor, 1e4, (double)maxIter, 4, 1e-1};
  std::vector<feNorm *> norms = {};
  TimeIntegrator       *solver;
  P->sys->_tAsmMat = P->sys->_tAsmRes = P->sys->_tSolve = 0.;
  P->sys->_nAsmMat = P->sys->_nAsmRes = P->sys->_nSolve = 0;
  if(createTimeIntegrator(solver, timeIntegratorScheme::STATIONARY, NL, P->sys, P->sol, P->mesh, norms,
                          {nullptr, 1, ""}) != FE_STATUS_OK)
    return -1;
  if(solver->makeSteps(1) != FE_STATUS_OK) {
    delete solver;
    return -2;
  }
  delete solver;
  std::memcpy(sol_out, P->sol->getSolution().data(), P->sol->getNumDOFs() * sizeof(double));
  out[0] = out[1] = 0.;
  if(P->uExact) {
    feNorm *eU = nullptr, *eP = nullptr;
    createNorm(eU, VECTOR_L2_ERROR, {P->uSpace}, P->sol, nullptr, P->uExact);
    createNorm(eP, L2_ERROR, {P->pSpace}, P->sol, P->pExact);
    out[0] = eU->compute();
    out[1] = eP->compute();
    delete eU;
    delete eP;
  } else if(P->sExact) {
    feNorm *eU = nullptr;
    createNorm(eU, L2_ERROR, {P->uSpace}, P->sol, P->sExact);
    out[0] = eU->compute();
    delete eU;
  }
  out[2] = P->sys->_tAsmMat;
  out[3] = P->sys->_tAsmRes;
  out[4] = P->sys->_tSolve;
  out[5] = P->sys->_nAsmMat;
  out[6] = P->sys->_nAsmRes;
  out[7] = P->sys->_nSolve;
  return 0;
}

// Transient run through the reference's unmodified time integrators (BDF1 / BDF2, src/feTimeIntegration.cpp) and Newton loop,
// from the recipe's initial state: scheme 1 = BDF1, 2 = BDF2; backend 0 = CPU stub (Eigen SparseLU), 1 = the product's
// feLinearSystemB200 (needs WITH_B200_ADAPTER).  opts = {pc, restart, linear max_iter}.  out: number of linear solves, total
// Krylov iterations, converged flag of the last solve.
int ref_transient(void *h, int backend, int scheme, double t0, double t1, int nSteps, double tolRes, double tolCor, int maxIter,
                  double relTol, const int *opts, double *sol_out, double *out)
{
  RefProblem *P = (RefProblem *)h;
  P->sol->initialize(P->mesh);
  feLinearSystem *sys = nullptr;
  bool            own = false;
  if(backend == 1) {
#ifdef WITH_B200_ADAPTER
    feB200Options o;
    o.preconditioner = opts[0];
    o.restart        = opts[1];
    if(createLinearSystemB200(sys, P->forms, P->numbering, o) != FE_STATUS_OK) return -1;
    sys->setRelativeTol(relTol);
    sys->setMaxIter(opts[2]);
    own = true;
#else
    return -9;
#endif
  } else
    sys = P->sys;
  feNLSolverOptions     NL{tolRes, tolCor, 1e4, (double)maxIter, 4, 1e-1};
  std::vector<feNorm *> norms = {};
  TimeIntegrator       *solver;
  const timeIntegratorScheme sch = scheme == 1 ? timeIntegratorScheme::BDF1 : timeIntegratorScheme::BDF2;
  if(createTimeIntegrator(solver, sch, NL, sys, P->sol, P->mesh, norms, {nullptr, 1, ""}, t0, t1, nSteps) != FE_STATUS_OK) {
    if(own) delete sys;
    return -2;
  }
  const feStatus st = solver->makeSteps(nSteps);
  delete solver;
  out[0] = out[1] = out[2] = 0.;
#ifdef WITH_B200_ADAPTER
  if(backend == 1) {
    feLinearSystemB200 *b = static_cast<feLinearSystemB200 *>(sys);
    out[0] = b->getNumSolves();
    out[1] = (double)b->getTotalKrylovIterations();
    out[2] = b->getLastSolveInfo().converged;
  }
#endif
  if(own) delete sys;
  if(st != FE_STATUS_OK) return -3;
  std::memcpy(sol_out, P->sol->getSolution().data(), P->sol->getNumDOFs() * sizeof(double));
  return 0;
}

#ifdef WITH_B200_ADAPTER
// The same stationary solve as ref_newton, but the UNMODIFIED createTimeIntegrator -> solveNewtonRaphson drives the
// product's feLinearSystemB200 backend (adapter/feLinearSystemB200.h -> libfeng_b200.so) instead of the CPU stub:
// the one-line change of tests/withLinearSolver/navier_stokes.cpp:101-108.
// opts: pc, restart, linear max_iter, scatter mode, device pattern (0/1).  out: errU, errP, nSolves, total Krylov
// iterations, last |Ax-b|, converged flag of the last solve.
static double g_b200_abs_tol = -1.; // < 0: keep feLinearSystem's default (1e-14, src/feLinearSystem.h:67)
void ref_set_b200_abs_tol(double v) { g_b200_abs_tol = v; }

int ref_newton_b200(void *h, double tolRes, double tolCor, int maxIter, double relTol, const int *opts, double *sol_out,
                    double *out)
{
  RefProblem *P = (RefProblem *)h;
  P->sol->initialize(P->mesh);
  feB200Options o;
  o.preconditioner = opts[0];
  o.restart        = opts[1];
  o.scatter        = opts[3];
  o.devicePattern  = opts[4] != 0;
  feLinearSystem *sys = nullptr;
  if(createLinearSystemB200(sys, P->forms, P->numbering, o) != FE_STATUS_OK) return -1;
  sys->setRelativeTol(relTol);
  if(g_b200_abs_tol >= 0.) sys->setAbsoluteTol(g_b200_abs_tol);
  sys->setMaxIter(opts[2]);
  feNLSolverOptions     NL{tolRes, tolCor, 1e4, (double)maxIter, 4, 1e-1};
  std::vector<feNorm *> norms = {};
  TimeIntegrator       *solver;
  if(createTimeIntegrator(solver, timeIntegratorScheme::STATIONARY, NL, sys, P->sol, P->mesh, norms, {nullptr, 1, ""}) !=
     FE_STATUS_OK) {
    delete sys;
    return -2;
  }
  const feStatus st = solver->makeSteps(1);
  delete solver;
  feLinearSystemB200 *b = static_cast<feLinearSystemB200 *>(sys);
  out[2] = b->getNumSolves();
  out[3] = (double)b->getTotalKrylovIterations();
  out[4] = b->getLastSolveInfo().norm_axb;
  out[5] = b->getLastSolveInfo().converged;
  delete sys;
  if(st != FE_STATUS_OK) return -3;
  std::memcpy(sol_out, P->sol->getSolution().data(), P->sol->getNumDOFs() * sizeof(double));
  return ref_error_norms(h, sol_out, out);
}
#endif

#ifdef WITH_B200_ADAPTER
// One assembly of the CURRENT state through the product's C++ adapter (feLinearSystemB200 built from the unmodified host
// objects: forms, spaces, numbering) -> CSR values and rhs of the CUDA backend.  what: bit 0 residual, bit 1 matrix.
int ref_assemble_b200(void *h, int what, int devicePattern, double *values, double *rhs)
{
  RefProblem   *P = (RefProblem *)h;
  feB200Options o;
  o.devicePattern = devicePattern != 0;
  feLinearSystem *sys = nullptr;
  if(createLinearSystemB200(sys, P->forms, P->numbering, o) != FE_STATUS_OK) return -1;
  solAtTimeN = P->stateN();
  sys->setToZero();
  if(what & 1) sys->assembleResiduals(P->sol);
  if(what & 2) sys->assembleMatrices(P->sol, false);
  feLinearSystemB200 *b  = static_cast<feLinearSystemB200 *>(sys);
  int                 rc = 0;
  if(values && b200_get_matrix_values(b->getHandle(), values) != B200_OK) rc = -2;
  if(rhs && b200_get_rhs(b->getHandle(), rhs) != B200_OK) rc = -3;
  if(b->getStatus() != FE_STATUS_OK) rc = -4;
  delete sys;
  return rc;
}
#endif

#ifdef WITH_B200_ADAPTER
// assemble + constrainEssentialComponents + applyPeriodicity of the CURRENT state through the adapter: the constrained system
// of the CUDA backend, to be compared entry by entry with ref_constrain (same pattern, same numbering)
int ref_constrain_b200(void *h, int devicePattern, double *values, double *rhs)
{
  RefProblem   *P = (RefProblem *)h;
  feB200Options o;
  o.devicePattern = devicePattern != 0;
  feLinearSystem *sys = nullptr;
  if(createLinearSystemB200(sys, P->forms, P->numbering, o) != FE_STATUS_OK) return -1;
  solAtTimeN = P->stateN();
  sys->setToZero();
  sys->assembleResiduals(P->sol);
  sys->assembleMatrices(P->sol, false);
  sys->constrainEssentialComponents(P->sol);
  sys->applyPeriodicity();
  feLinearSystemB200 *b  = static_cast<feLinearSystemB200 *>(sys);
  int                 rc = 0;
  int64_t             n = 0, nnz = 0;
  b200_get_pattern_size(b->getHandle(), &n, &nnz);
  if(nnz != P->sys->_nnz) rc = -5; // the device pattern must have the periodic extras of the reference's
  if(rc == 0 && b200_get_matrix_values(b->getHandle(), values) != B200_OK) rc = -2;
  if(rc == 0 && b200_get_rhs(b->getHandle(), rhs) != B200_OK) rc = -3;
  if(b->getStatus() != FE_STATUS_OK) rc = -4;
  delete sys;
  return rc;
}
#endif

// L2 error norms of an arbitrary nDOF solution vector against the recipe's analytic fields (feNorm).
int ref_error_norms(void *h, const double *sol, double *out)
{
  RefProblem         *P    = (RefProblem *)h;
  std::vector<double> save = P->sol->getSolution();
  std::memcpy(P->sol->getSolution().data(), sol, save.size() * sizeof(double));
  out[0] = out[1] = 0.;
  if(P->uExact) {
    feNorm *eU = nullptr, *eP = nullptr;
    createNorm(eU, VECTOR_L2_ERROR, {P->uSpace}, P->sol, nullptr, P->uExact);
    createNorm(eP, L2_ERROR, {P->pSpace}, P->sol, P->pExact);
    out[0] = eU->compute();
    out[1] = eP->compute();
    delete eU;
    delete eP;
  } else if(P->sExact) {
    feNorm *eU = nullptr;
    createNorm(eU, L2_ERROR, {P->uSpace}, P->sol, P->sExact);
    out[0] = eU->compute();
    delete eU;
  }
  P->sol->getSolution() = save;
  return 0;
}

// periodic (master, slave) DOF pairs of the numbering (feMetaNumber::PeriodicDOF, src/feNumber.h:234)
int ref_periodic_pairs(void *h, int64_t *master, int64_t *slave, int64_t *n)
{
  RefProblem *P = (RefProblem *)h;
  int64_t     k = 0;
  for(const auto &pr : P->numbering->PeriodicDOF()) {
    if(master) master[k] = pr.first;
    if(slave) slave[k] = pr.second;
    ++k;
  }
  *n = k;
  return 0;
}

int ref_constraint_rows(void *h, int64_t *rows, int64_t *n)
{
  RefProblem *P = (RefProblem *)h;
  if(!P->sys->_constraintInit) P->sys->initConstraintRows(P->sol);
  *n = (int64_t)P->sys->_rowsToConstrain.size();
  if(rows)
    for(size_t i = 0; i < P->sys->_rowsToConstrain.size(); ++i) rows[i] = P->sys->_rowsToConstrain[i];
  return 0;
}

} // extern "C"
