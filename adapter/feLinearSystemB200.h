// feLinearSystemB200 -- a new feLinearSystem backend for arthurbawin/feNG, next to feLinearSystemPETSc
// (src/feLinearSystem.h:174) and feLinearSystemMklPardiso (src/feLinearSystem.h:276).
//
// Header-only.  Compile against the UNMODIFIED reference headers (-I<feNG>/src) and link libfeng_b200.so
// (include/feng_b200.h).  A driver changes one line: where it called
//     createLinearSystem(system, MKLPARDISO, forms, &numbering)      (e.g. tests/withLinearSolver/navier_stokes.cpp:101-108)
// it calls
//     createLinearSystemB200(system, forms, &numbering)
// and the rest -- createTimeIntegrator, solveNewtonRaphson (src/feNonLinearSolver.cpp:38-182), norms -- is untouched.
//
// What the constructor extracts ONCE from the host objects (all through public API, except the two documented peeks):
//   vertices, connectivity           feMesh::getVertices, feCncGeo::getVertexConnectivity   (src/feMesh.h:130, src/feCncGeo.h:191)
//   element -> DOF tables            feSpace::initializeAddressingVector                      (src/feSpace.h:454)
//   scalar basis tables, weights     feSpace::getFunctionAtQuadNode / getd...AtQuadNode / L() (src/feSpace.h:381-408, :446-450)
//   nInc, nDOF, periodic pairs       feMetaNumber                                             (src/feNumber.h:202-203, :234)
//   sparsity pattern                 feEZCompressedRowStorage through a one-line derived struct exposing the protected
//                                    ia_Pardiso/ja_Pardiso (src/feCompressedRowStorage.h:33-37)
//   weak-form kind                   feBilinearForm::getID                                    (src/feBilinearForm.h:165)
//   weak-form coefficients           protected members of the feSysElm_* classes (src/feSysElm.h:268,311,584,650,689,
//                                    751-752,887,976-977,1025,1144), read through pointer-to-member of a derived struct;
//                                    callbacks are sampled and must be constant in space (else FE_STATUS_ERROR), source
//                                    callbacks are tabulated at every (element, quadrature node) and re-tabulated when the
//                                    time changes.
// Per Newton iteration only the state vectors cross the boundary (feSolution::getSolution / getSolutionDot, c0, tn).
#ifndef _FELINEARSYSTEMB200_
#define _FELINEARSYSTEMB200_

#include "feLinearSystem.h"
#include "feSysElm.h"
#include "feSpace.h"
#include "feCncGeo.h"
#include "feMesh.h"

#include "feng_b200.h"

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <cstdio>
#include <fstream>
#include <map>

struct feB200Options {
  int device       = 0;
  int preconditioner = B200_PC_AUTO;   // B200_PC_*: Schur/multigrid for Taylor-Hood, multigrid for scalar systems, Jacobi for CHNS
  int restart      = 30;               // GMRES restart length (PETSc default)
  int scatter      = B200_SCATTER_ATOMIC;
  bool devicePattern = false;          // build the EZCRS pattern on the device instead of taking the host one
  bool alwaysUpload  = false;          // re-upload the state in assembleMatrices even right after assembleResiduals
  // keep the state resident on the device.  After the first upload the device copy -- which correctSolution
  // updates in place -- is authoritative for the UNKNOWNS; every later assemble only sends the essential-DOF tail of the host
  // vector (entries >= nInc, rewritten by feSolution::initializeEssentialBC each time step, src/feTimeIntegration.cpp:523-536).
  // Only valid for drivers that never write unknowns of feSolution between two assemblies other than through correctSolution
  // (solveNewtonRaphson on a stationary problem does not); transient drivers keep it off because BDFContainer rewrites solDot
  // on the host (src/feSolutionContainer.cpp:338-348) -- a B200-aware host calls b200_state_push / b200_state_bdf instead.
  bool deviceResidentState = false;
};

extern std::vector<double> solAtTimeN; // src/feNonLinearSolver.cpp:35

namespace feB200detail
{
  struct FormPeek : public feBilinearForm {
    static feSysElm *sysElm(feBilinearForm *f) { return f->*(&FormPeek::_sysElm); }
  };
  struct PatternPeek : public feEZCompressedRowStorage {
    using feEZCompressedRowStorage::feEZCompressedRowStorage;
    const std::vector<feInt> &ia() const { return ia_Pardiso; }
    const std::vector<feInt> &ja() const { return ja_Pardiso; }
  };
  template <class T> struct SysPeek : public T {
    static const feFunction *coeff(const T *s) { return s->*(&SysPeek::_coeff); }
  };
  template <class T> struct SourcePeek : public T {
    static auto source(const T *s) -> decltype(s->*(&SourcePeek::_source)) { return s->*(&SourcePeek::_source); }
  };
  template <class T> struct ViscPeek : public T {
    static const feFunction *visc(const T *s) { return s->*(&ViscPeek::_viscosity); }
  };
  template <class T> struct DiffPeek : public T {
    static const feFunction *diff(const T *s) { return s->*(&DiffPeek::_diffusivity); }
  };
  // CHNS_Abels<2> / CHNS_MassAveraged<2> keep their property callbacks and constants protected
  // (src/feSysElm.h:1272-1284, :1355-1368)
  template <class C> struct ChnsPeekT : public C {
    static const feFunction *density(const C *s) { return s->*(&ChnsPeekT::_density); }
    static const feFunction *drhodphi(const C *s) { return s->*(&ChnsPeekT::_drhodphi); }
    static const feFunction *viscosity(const C *s) { return s->*(&ChnsPeekT::_viscosity); }
    static const feFunction *mobility(const C *s) { return s->*(&ChnsPeekT::_mobility); }
    static const feVectorFunction *volumeForce(const C *s) { return s->*(&ChnsPeekT::_volumeForce); }
    static const feVectorFunction *sourceU(const C *s) { return s->*(&ChnsPeekT::_sourceU); }
    static const feFunction *sourceP(const C *s) { return s->*(&ChnsPeekT::_sourceP); }
    static const feFunction *sourcePhi(const C *s) { return s->*(&ChnsPeekT::_sourcePhi); }
    static const feFunction *sourceMu(const C *s) { return s->*(&ChnsPeekT::_sourceMu); }
    static double surfaceTension(const C *s) { return s->*(&ChnsPeekT::_surfaceTension); }
    static double epsilon(const C *s) { return s->*(&ChnsPeekT::_epsilon); }
  };
  struct MassAveragedPeek : public CHNS_MassAveraged<2> {
    static double alpha(const CHNS_MassAveraged<2> *s) { return s->*(&MassAveragedPeek::_alpha); }
  };
  // CHNS_Khanwale<2> has no surface tension / epsilon members: Re, Pe, Cn, We, Fr, rhoA, rhoB (src/feSysElm.h:1449)
  struct KhanwalePeek : public CHNS_Khanwale<2> {
    using C = CHNS_Khanwale<2>;
    static const feFunction *density(const C *s) { return s->*(&KhanwalePeek::_density); }
    static const feFunction *drhodphi(const C *s) { return s->*(&KhanwalePeek::_drhodphi); }
    static const feFunction *viscosity(const C *s) { return s->*(&KhanwalePeek::_viscosity); }
    static const feFunction *mobility(const C *s) { return s->*(&KhanwalePeek::_mobility); }
    static const feVectorFunction *volumeForce(const C *s) { return s->*(&KhanwalePeek::_volumeForce); }
    static const feVectorFunction *sourceU(const C *s) { return s->*(&KhanwalePeek::_sourceU); }
    static const feFunction *sourceP(const C *s) { return s->*(&KhanwalePeek::_sourceP); }
    static const feFunction *sourcePhi(const C *s) { return s->*(&KhanwalePeek::_sourcePhi); }
    static const feFunction *sourceMu(const C *s) { return s->*(&KhanwalePeek::_sourceMu); }
    static double surfaceTension(const C *) { return 0.; }
    static double epsilon(const C *) { return 1.; }
    static void numbers(const C *s, double *out)
    {
      out[0] = s->*(&KhanwalePeek::_Re);
      out[1] = s->*(&KhanwalePeek::_Pe);
      out[2] = s->*(&KhanwalePeek::_Cn);
      out[3] = s->*(&KhanwalePeek::_We);
      out[4] = s->*(&KhanwalePeek::_Fr);
      out[5] = s->*(&KhanwalePeek::_rhoA);
      out[6] = s->*(&KhanwalePeek::_rhoB);
    }
  };
  template <class C> struct ChnsPeekOf { using type = ChnsPeekT<C>; };
  template <> struct ChnsPeekOf<CHNS_Khanwale<2>> { using type = KhanwalePeek; };
} // namespace feB200detail

class feLinearSystemB200 : public feLinearSystem
{
protected:
  b200_system  *_sys = nullptr;
  feB200Options _opt;
  feInt         _nInc = 0, _nDOF = 0;
  feStatus      _status = FE_STATUS_OK;
  feMesh       *_mesh = nullptr;
  const feCncGeo *_cnc = nullptr;
  int           _dim = 0, _nElm = 0, _nQuad = 0;
  std::vector<feSpace *> _spaceList; // engine space id -> host space
  bool          _stateFresh = false;
  bool          _needSolutionN = false;
  bool          _constraintInit = false;
  bool          _deviceCurrent = false;
  int           _numUploads = 0;
  std::vector<int64_t> _essIdx;
  b200_solve_info _lastInfo{};
  int           _numSolves = 0;
  long          _totalKrylovIterations = 0;

  struct SourceForm {
    int                     formId;
    const feFunction       *scalar;
    const feVectorFunction *vector;
    feSpace                *geoSpace;
    int                     cncGeoTag;
    double                  time;
  };
  std::vector<SourceForm> _sources;

  bool fail(const std::string &what)
  {
    feErrorMsg(FE_STATUS_ERROR, "feLinearSystemB200: %s: %s", what.c_str(), b200_last_error());
    _status = FE_STATUS_ERROR;
    return false;
  }
  bool ok(int rc, const char *what) { return rc >= 0 ? true : fail(what); }

  int spaceId(feSpace *s)
  {
    for(size_t i = 0; i < _spaceList.size(); ++i)
      if(_spaceList[i] == s) return (int)i;
    // scalar Lagrange tables of the space: for a vector space function a*nc+c is phi_a e_c (src/feSpace_2D.cpp:41-55),
    // so the scalar basis is component c=0 of functions 0, nc, 2nc, ...
    const int nF = s->getNumFunctions(), nC = s->getNumComponents(), nS = nF / nC;
    std::vector<int32_t> adr((size_t)_nElm * nF);
    std::vector<feInt>   a(nF);
    for(int e = 0; e < _nElm; ++e) {
      s->initializeAddressingVector(e, a);
      for(int j = 0; j < nF; ++j) adr[(size_t)nF * e + j] = (int32_t)a[j];
    }
    std::vector<double> L((size_t)_nQuad * nS), dL((size_t)_nQuad * nS * _dim);
    const std::vector<double> &r = s->getRQuadraturePoints(), &ss = s->getSQuadraturePoints(), &t = s->getTQuadraturePoints();
    for(int k = 0; k < _nQuad; ++k) {
      if(nC == 1) {
        for(int i = 0; i < nS; ++i) {
          L[(size_t)k * nS + i]                = s->getFunctionAtQuadNode(i, k);
          dL[((size_t)k * nS + i) * _dim + 0] = s->getdFunctiondrAtQuadNode(i, k);
          if(_dim >= 2) dL[((size_t)k * nS + i) * _dim + 1] = s->getdFunctiondsAtQuadNode(i, k);
          if(_dim >= 3) dL[((size_t)k * nS + i) * _dim + 2] = s->getdFunctiondtAtQuadNode(i, k);
        }
      } else {
        double              rr[3] = {r[k], ss[k], t[k]};
        std::vector<double> l = s->L(rr), da = s->dLdr(rr), db = s->dLds(rr);
        // vector tables are [i][c]; take component 0 of functions i = a*nC
        for(int i = 0; i < nS; ++i) {
          L[(size_t)k * nS + i]                = l[(size_t)(i * nC) * nC + 0];
          dL[((size_t)k * nS + i) * _dim + 0] = da[(size_t)(i * nC) * nC + 0];
          dL[((size_t)k * nS + i) * _dim + 1] = db[(size_t)(i * nC) * nC + 0];
        }
      }
    }
    const int id = b200_add_space(_sys, nS, nC, adr.data(), L.data(), dL.data());
    if(id < 0) {
      fail("b200_add_space");
      return -1;
    }
    _spaceList.push_back(s);
    return id;
  }

  // value of a coefficient callback.  The fused kernels take constants (include/feng_b200.h, b200_add_form): the callback is
  // evaluated at EVERY (element, quadrature node) -- as the reference does on every element visit, src/feVectorSysElm.cpp:1183 --
  // and again at a second time and a second value of args.u; any variation is reported (FE_STATUS_ERROR) instead of being
  // frozen into one number.
  bool constantValue(const feFunction *f, feSpace *geoSpace, int cncGeoTag, double t, double &value)
  {
    std::vector<double> coord(3 * _cnc->getNumVerticesPerElem(), 0.); // feMesh::getCoord does not resize (src/feMesh.cpp:119-134)
    bool first = true;
    for(int pass = 0; pass < 2; ++pass) {
      feFunctionArguments args(pass == 0 ? t : t + 0.37);
      args.u = pass == 0 ? 0. : 0.61;
      // second pass (time / solution dependence): a sample of the elements is enough, space dependence was settled by the first
      const int step = pass == 0 ? 1 : std::max(1, _nElm / 16);
      for(int e = 0; e < _nElm; e += step) {
        _mesh->getCoord(cncGeoTag, e, coord);
        for(int k = 0; k < _nQuad; ++k) {
          geoSpace->interpolateVectorFieldAtQuadNode(coord, k, args.pos);
          const double v = f->eval(args);
          if(first) {
            value = v;
            first = false;
          } else if(std::fabs(v - value) > 1e-14 * std::max(1., std::fabs(value)))
            return false;
        }
      }
    }
    return true;
  }

  // Space- / time-dependent coefficient of a scalar diffusion form: tabulated at every (element, quadrature node), where the
  // reference evaluates it (src/feSysElm.cpp:538, :566), and re-tabulated when the time changes.  Solution-dependent callbacks
  // (args.u) are rejected: the engine's Jacobian has no term for them.
  struct CoeffForm {
    int               formId;
    const feFunction *f;
    feSpace          *geoSpace;
    int               cncGeoTag;
    double            time;
  };
  std::vector<CoeffForm> _coeffs;

  void tabulateCoefficient(const CoeffForm &cf, double t, double u, std::vector<double> &tab)
  {
    tab.resize((size_t)_nElm * _nQuad);
    feFunctionArguments args(t);
    args.u = u;
    std::vector<double> coord(3 * _cnc->getNumVerticesPerElem(), 0.);
    for(int e = 0; e < _nElm; ++e) {
      _mesh->getCoord(cf.cncGeoTag, e, coord);
      for(int k = 0; k < _nQuad; ++k) {
        cf.geoSpace->interpolateVectorFieldAtQuadNode(coord, k, args.pos);
        tab[(size_t)e * _nQuad + k] = cf.f->eval(args);
      }
    }
  }

  template <class T> bool addDiffusionForm(feBilinearForm *f, feSysElm *se, int kind, int su)
  {
    const T *s = dynamic_cast<const T *>(se);
    if(!s) return fail("weak form class does not match its elementSystemType");
    const feFunction *cb = feB200detail::SysPeek<T>::coeff(s);
    double c = 1.;
    if(constantValue(cb, f->_geoSpace, f->getCncGeoTag(), 0., c)) return ok(b200_add_form(_sys, kind, su, -1, c, 1., nullptr), "b200_add_form");
    CoeffForm           cf{-1, cb, f->_geoSpace, f->getCncGeoTag(), 0.};
    std::vector<double> t0, t1;
    tabulateCoefficient(cf, 0., 0., t0);
    tabulateCoefficient(cf, 0., 0.61, t1);
    for(size_t i = 0; i < t0.size(); ++i)
      if(std::fabs(t0[i] - t1[i]) > 1e-14 * std::max(1., std::fabs(t0[i])))
        return fail("solution-dependent diffusivity callbacks are not supported (the Jacobian would miss d k / d u)");
    cf.formId = b200_add_form(_sys, kind, su, -1, 1., 1., nullptr);
    if(cf.formId < 0) return fail("b200_add_form");
    if(!ok(b200_set_form_coefficient(_sys, cf.formId, t0.data()), "b200_set_form_coefficient")) return false;
    _coeffs.push_back(cf);
    return true;
  }

  void tabulateSource(const SourceForm &sf, double t, std::vector<double> &tab)
  {
    const int nc = sf.vector ? _dim : 1;
    tab.resize((size_t)_nElm * _nQuad * nc);
    feFunctionArguments args(t);
    std::vector<double> coord(3 * _cnc->getNumVerticesPerElem(), 0.), S(3, 0.);
    for(int e = 0; e < _nElm; ++e) {
      _mesh->getCoord(sf.cncGeoTag, e, coord);
      for(int k = 0; k < _nQuad; ++k) {
        sf.geoSpace->interpolateVectorFieldAtQuadNode(coord, k, args.pos); // as src/feVectorSysElm.cpp:118, src/feSysElm.cpp:20
        if(sf.vector) {
          (*sf.vector)(args, S);
          for(int c = 0; c < nc; ++c) tab[((size_t)e * _nQuad + k) * nc + c] = S[c];
        } else
          tab[(size_t)e * _nQuad + k] = sf.scalar->eval(args);
      }
    }
  }

  // The monolithic CHNS weak form (CHNS_Solver, src/CHNS_Solver.cpp:236-420).  Its property callbacks are functions of
  // the phase marker (args.u): the engine evaluates the laws CHNS_Solver installs (src/CHNS_Solver.cpp:124-235) itself,
  // so the callbacks are PROBED here and must reproduce one of those laws: linear mixing in phi (optionally clipped to
  // [-1, 1]) for density and viscosity, constant or degenerate mobility M |1 - phi^2|, constant force and sources.
  static double evalAt(const feFunction *f, double phi)
  {
    feFunctionArguments args(0.);
    args.u = phi;
    return f->eval(args);
  }
  bool probeLinearLaw(const feFunction *f, double &a, double &b, bool &limiter)
  {
    a = evalAt(f, 1.);
    b = evalAt(f, -1.);
    const double mid = evalAt(f, 0.3), out = evalAt(f, 1.7), scale = std::max(1., std::max(std::fabs(a), std::fabs(b)));
    if(std::fabs(mid - ((a - b) / 2. * 0.3 + (a + b) / 2.)) > 1e-13 * scale) return false;
    if(std::fabs(out - a) <= 1e-13 * scale)
      limiter = true;
    else if(std::fabs(out - ((a - b) / 2. * 1.7 + (a + b) / 2.)) <= 1e-13 * scale)
      limiter = false;
    else
      return false;
    return true;
  }
  bool constantVector(const feVectorFunction *f, double *out)
  {
    std::vector<double> r0(3, 0.), r1(3, 0.);
    feFunctionArguments a0(0.), a1(0.);
    a0.u = 0.2;
    a1.u = -0.9;
    a1.pos[0] = 0.37;
    a1.pos[1] = -0.61;
    (*f)(a0, r0);
    (*f)(a1, r1);
    for(int c = 0; c < 2; ++c) {
      if(std::fabs(r0[c] - r1[c]) > 1e-14 * std::max(1., std::fabs(r0[c]))) return false;
      out[c] = r0[c];
    }
    out[2] = 0.;
    return true;
  }
  template <class C> bool addChnsForm(feBilinearForm *f, feSysElm *se, int kind)
  {
    auto *s = dynamic_cast<const C *>(se);
    if(!s || f->_intSpaces.size() != 4) return fail("CHNS form must be a CHNS_Abels<2> / CHNS_MassAveraged<2> on {U, P, Phi, Mu}");
    using Pk = typename feB200detail::ChnsPeekOf<C>::type;
    b200_chns_params prm{};
    if constexpr(std::is_same<C, CHNS_MassAveraged<2>>::value) {
      prm.mass_alpha = feB200detail::MassAveragedPeek::alpha(s);
      _needSolutionN = true; // phi at the previous time step: the global solAtTimeN goes to the device with the state
    }
    if constexpr(std::is_same<C, CHNS_Khanwale<2>>::value) {
      feB200detail::KhanwalePeek::numbers(s, prm.khanwale);
      _needSolutionN = true; // every field at the previous time step and the time step
    }
    bool limRho = false, limVisc = false;
    if(!probeLinearLaw(Pk::density(s), prm.rho_a, prm.rho_b, limRho) || !probeLinearLaw(Pk::viscosity(s), prm.visc_a, prm.visc_b, limVisc) ||
       limRho != limVisc)
      return fail("CHNS density / viscosity callbacks are not the (clipped) linear mixing laws of CHNS_Solver");
    if(std::fabs(evalAt(Pk::drhodphi(s), 0.3) - (prm.rho_a - prm.rho_b) / 2.) > 1e-13 * std::max(1., std::fabs(prm.rho_a)))
      return fail("CHNS drhodphi callback does not match the density law");
    prm.limiter = limRho ? 1 : 0;
    {
      const feFunction *m = Pk::mobility(s);
      const double m0 = evalAt(m, 0.), m1 = evalAt(m, 1.), mh = evalAt(m, 0.5), mo = evalAt(m, 1.5);
      prm.mobility = m0;
      if(std::fabs(m1 - m0) <= 1e-14 * std::max(1., m0) && std::fabs(mh - m0) <= 1e-14 * std::max(1., m0))
        prm.degenerate_mobility = 0;
      else if(std::fabs(m1) <= 1e-14 * std::max(1., m0) && std::fabs(mh - 0.75 * m0) <= 1e-13 * std::max(1., m0) &&
              std::fabs(mo - 1.25 * m0) <= 1e-13 * std::max(1., m0))
        prm.degenerate_mobility = 1;
      else
        return fail("CHNS mobility callback is neither constant nor M |1 - phi^2|");
    }
    prm.surface_tension = Pk::surfaceTension(s);
    prm.epsilon         = Pk::epsilon(s);
    if(!constantVector(Pk::volumeForce(s), prm.force) || !constantVector(Pk::sourceU(s), prm.source_u))
      return fail("CHNS volume force / momentum source must be constant");
    if(!constantValue(Pk::sourceP(s), f->_geoSpace, f->getCncGeoTag(), 0., prm.source_p) ||
       !constantValue(Pk::sourcePhi(s), f->_geoSpace, f->getCncGeoTag(), 0., prm.source_phi) ||
       !constantValue(Pk::sourceMu(s), f->_geoSpace, f->getCncGeoTag(), 0., prm.source_mu))
      return fail("CHNS scalar sources must be constant");
    int sp[4];
    for(int k = 0; k < 4; ++k)
      if((sp[k] = spaceId(f->_intSpaces[k])) < 0) return false;
    return ok(b200_add_form_chns(_sys, kind, sp[0], sp[1], sp[2], sp[3], &prm), "b200_add_form_chns");
  }

  template <class T> bool addCoeffForm(feBilinearForm *f, feSysElm *se, int kind, int su, int sp, const feFunction *param)
  {
    const T *s = dynamic_cast<const T *>(se);
    if(!s) return fail("weak form class does not match its elementSystemType");
    double c = 1., p = 1.;
    if(!constantValue(feB200detail::SysPeek<T>::coeff(s), f->_geoSpace, f->getCncGeoTag(), 0., c))
      return fail("non-constant coefficient callbacks are not supported by the fused kernels");
    if(param && !constantValue(param, f->_geoSpace, f->getCncGeoTag(), 0., p))
      return fail("non-constant viscosity/diffusivity callbacks are not supported by the fused kernels");
    return ok(b200_add_form(_sys, kind, su, sp, c, p, nullptr), "b200_add_form");
  }

  template <int dim> bool addFormDim(feBilinearForm *f, feSysElm *se)
  {
    const int id = (int)f->getID();
    const int s0 = spaceId(f->_intSpaces[0]);
    const int s1 = f->_intSpaces.size() > 1 ? spaceId(f->_intSpaces[1]) : -1;
    if(s0 < 0) return false;
    switch(id) {
      case VECTOR_CONVECTIVE_ACCELERATION:
        return addCoeffForm<feSysElm_VectorConvectiveAcceleration<dim>>(f, se, id, s0, -1, nullptr);
      case DIV_NEWTONIAN_STRESS: {
        auto *s = dynamic_cast<const feSysElm_DivergenceNewtonianStress<dim> *>(se);
        if(!s) return fail("DivergenceNewtonianStress cast");
        return addCoeffForm<feSysElm_DivergenceNewtonianStress<dim>>(
          f, se, id, s0, s1, feB200detail::ViscPeek<feSysElm_DivergenceNewtonianStress<dim>>::visc(s));
      }
      case MIXED_DIVERGENCE: // declared on {p, u}: _intSpaces[0] = P (tests/withLinearSolver/navier_stokes.cpp:85)
        return addCoeffForm<feSysElm_MixedDivergence<dim>>(f, se, id, s0, s1, nullptr);
      case MIXED_GRADIENT: return addCoeffForm<feSysElm_MixedGradient<dim>>(f, se, id, s0, s1, nullptr);
      case VECTOR_DIFFUSION: {
        auto *s = dynamic_cast<const feSysElm_VectorDiffusion<dim> *>(se);
        if(!s) return fail("VectorDiffusion cast");
        return addCoeffForm<feSysElm_VectorDiffusion<dim>>(f, se, id, s0, -1,
                                                           feB200detail::DiffPeek<feSysElm_VectorDiffusion<dim>>::diff(s));
      }
      case TRANSIENT_VECTOR_MASS: return addCoeffForm<feSysElm_TransientVectorMass<dim>>(f, se, id, s0, -1, nullptr);
      case VECTOR_SOURCE: {
        auto *s = dynamic_cast<const feSysElm_VectorSource<dim> *>(se);
        if(!s) return fail("VectorSource cast");
        SourceForm sf{-1, nullptr, feB200detail::SourcePeek<feSysElm_VectorSource<dim>>::source(s), f->_geoSpace, f->getCncGeoTag(), 0.};
        std::vector<double> tab;
        tabulateSource(sf, 0., tab);
        sf.formId = b200_add_form(_sys, id, s0, -1, 1., 0., tab.data());
        if(sf.formId < 0) return fail("b200_add_form(source)");
        _sources.push_back(sf);
        return true;
      }
      default: break;
    }
    return fail("weak form " + f->getWeakFormName() + " is not supported by the B200 backend");
  }

  bool addForm(feBilinearForm *f)
  {
    feSysElm *se = feB200detail::FormPeek::sysElm(f);
    const int id = (int)f->getID();
    const int s0 = spaceId(f->_intSpaces[0]);
    if(s0 < 0) return false;
    switch(id) {
      case SOURCE: {
        auto *s = dynamic_cast<const feSysElm_Source *>(se);
        if(!s) return fail("Source cast");
        SourceForm sf{-1, feB200detail::SourcePeek<feSysElm_Source>::source(s), nullptr, f->_geoSpace, f->getCncGeoTag(), 0.};
        std::vector<double> tab;
        tabulateSource(sf, 0., tab);
        sf.formId = b200_add_form(_sys, id, s0, -1, 1., 0., tab.data());
        if(sf.formId < 0) return fail("b200_add_form(source)");
        _sources.push_back(sf);
        return true;
      }
      case CHNS_ABELS: return addChnsForm<CHNS_Abels<2>>(f, se, B200_FORM_CHNS_ABELS);
      case CHNS_MASS_AVERAGED: return addChnsForm<CHNS_MassAveraged<2>>(f, se, B200_FORM_CHNS_MASS_AVERAGED);
      case CHNS_KHANWALE: return addChnsForm<CHNS_Khanwale<2>>(f, se, B200_FORM_CHNS_KHANWALE);
      case TRANSIENT_MASS: return addCoeffForm<feSysElm_TransientMass>(f, se, id, s0, -1, nullptr);
      case DIFFUSION:
        // diffusivity is the form's only callback: kind DIFFUSION uses coeff x param with param = 1
        if(_dim == 2) return addDiffusionForm<feSysElm_Diffusion<2>>(f, se, id, s0);
        return addDiffusionForm<feSysElm_Diffusion<3>>(f, se, id, s0);
      default: break;
    }
    if(_dim == 2) return addFormDim<2>(f, se);
    return fail("vector-valued weak forms exist only for dim = 2 in the reference (src/feVectorSysElm.cpp:1246,1536)");
  }

  void refreshSources(const feSolution *sol)
  {
    const double t = sol->getCurrentTime();
    std::vector<double> tab;
    for(auto &sf : _sources) {
      if(sf.time == t) continue;
      tabulateSource(sf, t, tab);
      ok(b200_set_source(_sys, sf.formId, tab.data()), "b200_set_source");
      sf.time = t;
    }
    for(auto &cf : _coeffs) {
      if(cf.time == t) continue;
      tabulateCoefficient(cf, t, 0., tab);
      ok(b200_set_form_coefficient(_sys, cf.formId, tab.data()), "b200_set_form_coefficient");
      cf.time = t;
    }
  }

  void upload(const feSolution *sol)
  {
    refreshSources(sol);
    if(_opt.deviceResidentState && _deviceCurrent && sol->getC0() == 0.) {
      if(_essIdx.empty())
        for(feInt i = _nInc; i < _nDOF; ++i) _essIdx.push_back((int64_t)i);
      if(!_essIdx.empty())
        ok(b200_set_essential(_sys, (int64_t)_essIdx.size(), _essIdx.data(), sol->getSolution().data() + _nInc), "b200_set_essential");
      return;
    }
    _deviceCurrent = true;
    ++_numUploads;
    ok(b200_set_solution(_sys, sol->getSolution().data(), sol->getSolutionDot().data(), sol->getC0(), sol->getCurrentTime()),
       "b200_set_solution");
    // feBilinearForm::initialize reads the previous time step from the global solAtTimeN (src/feBilinearForm.cpp:277,347)
    if(_needSolutionN)
      ok(b200_set_solution_n(_sys, (feInt)solAtTimeN.size() == _nDOF ? solAtTimeN.data() : nullptr, sol->getTimeStep()),
         "b200_set_solution_n");
  }

  // rows of essential vector components, as src/feLinearSystemMklPardiso.cpp:998-1041
  void initConstraints(const feSolution *sol)
  {
    _constraintInit = true;
    std::vector<int64_t> rows;
    std::vector<feInt>   adr;
    for(const auto &space : sol->_spaces) {
      const int nComponents = space->getNumComponents();
      if(nComponents <= 1) continue;
      const int nFunctions = space->getNumFunctions(), nElm = space->getNumElements();
      adr.resize(nFunctions);
      for(int c = 0; c < nComponents; ++c) {
        if(!space->isEssentialComponent(c)) continue;
        for(int e = 0; e < nElm; ++e) {
          space->initializeAddressingVector(e, adr);
          for(int j = 0; j < nFunctions; ++j)
            if(j % nComponents == c && adr[j] < _nInc) rows.push_back(adr[j]);
        }
      }
    }
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    if(!rows.empty()) ok(b200_set_constraints(_sys, (int64_t)rows.size(), rows.data(), 0, nullptr, nullptr), "b200_set_constraints");
  }

public:
  feLinearSystemB200(const std::vector<feBilinearForm *> forms, const feMetaNumber *numbering, const feB200Options &opt = feB200Options())
    : feLinearSystem(forms, numbering), _opt(opt)
  {
    _recomputeMatrix = true;
    _nInc            = numbering->getNbUnknowns();
    _nDOF            = numbering->getNbDOFs();
    if(forms.empty()) {
      fail("no weak form");
      return;
    }
    if(!ok(b200_create(&_sys, opt.device), "b200_create")) return;
    // one interior connectivity for all forms
    _cnc = forms[0]->getCncGeo();
    for(auto *f : forms)
      if(f->getCncGeo() != _cnc) {
        fail("all weak forms must live on the same geometric connectivity");
        return;
      }
    feSpace *s0 = forms[0]->_intSpaces[0];
    _mesh       = s0->getMeshPtr();
    _dim        = _cnc->getDim();
    _nElm       = _cnc->getNumElements();
    _nQuad      = s0->getNumQuadPoints();
    const int nv = _cnc->getNumVerticesPerElem();
    {
      auto               &V = _mesh->getVertices();
      std::vector<double> xyz(3 * V.size());
      for(size_t i = 0; i < V.size(); ++i) {
        xyz[3 * i + 0] = V[i].x();
        xyz[3 * i + 1] = V[i].y();
        xyz[3 * i + 2] = V[i].z();
      }
      std::vector<int32_t> conn((size_t)_nElm * nv);
      for(int e = 0; e < _nElm; ++e)
        for(int j = 0; j < nv; ++j) conn[(size_t)nv * e + j] = _cnc->getVertexConnectivity(e, j);
      if(!ok(b200_set_mesh(_sys, _dim, (int64_t)V.size(), xyz.data(), _nElm, nv, conn.data()), "b200_set_mesh")) return;
    }
    if(!ok(b200_set_quadrature(_sys, _nQuad, s0->getQuadratureWeights().data()), "b200_set_quadrature")) return;
    // register spaces in the order the forms use them, then the pattern, then the forms
    for(auto *f : forms)
      for(auto *s : f->_intSpaces)
        if(spaceId(s) < 0) return;
    {
      // periodic pairs first: the device pattern builder adds their (slave, master) entries (src/feCompressedRowStorage.cpp:96-107)
      std::vector<int64_t> master, slave;
      for(const auto &pr : numbering->PeriodicDOF()) {
        master.push_back(pr.first);
        slave.push_back(pr.second);
      }
      if(!master.empty() && !ok(b200_set_periodic(_sys, (int64_t)master.size(), master.data(), slave.data()), "b200_set_periodic")) return;
    }
    if(!opt.devicePattern) {
      feB200detail::PatternPeek crs((int)_nInc, _formMatrices, _numMatrixForms, numbering);
      std::vector<int64_t> ia(crs.ia().begin(), crs.ia().end());
      std::vector<int32_t> ja(crs.ja().begin(), crs.ja().end());
      if(!ok(b200_set_pattern(_sys, _nInc, _nDOF, ia.data(), ja.data()), "b200_set_pattern")) return;
    }
    for(auto *f : forms)
      if(!addForm(f)) return;
    if(opt.devicePattern && !ok(b200_build_pattern(_sys, _nInc, _nDOF), "b200_build_pattern")) return;
    if(opt.scatter == B200_SCATTER_COLORED) {
      const std::vector<int> &c = _cnc->getColorElm();
      std::vector<int32_t>    col(c.begin(), c.end());
      if(!ok(b200_set_colors(_sys, _cnc->getNbColor(), col.data()), "b200_set_colors")) return;
      if(!ok(b200_set_scatter_mode(_sys, B200_SCATTER_COLORED), "b200_set_scatter_mode")) return;
    }
    ok(b200_finalize(_sys), "b200_finalize");
  }

  ~feLinearSystemB200() { b200_destroy(_sys); }

  feStatus getStatus() const { return _status; }
  const b200_solve_info &getLastSolveInfo() const { return _lastInfo; }
  int  getNumSolves() const { return _numSolves; }
  int  getNumStateUploads() const { return _numUploads; }
  long getTotalKrylovIterations() const { return _totalKrylovIterations; }
  b200_system *getHandle() { return _sys; }

  feInt getSystemSize() const { return _nInc; }
  void  getRHSMaxNorm(double *norm) const { b200_rhs_max_norm(_sys, norm); }
  void  getResidualMaxNorm(double *norm) const { b200_du_max_norm(_sys, norm); }

  void setToZero()
  {
    ok(b200_set_to_zero(_sys, _recomputeMatrix ? 3 : 1), "b200_set_to_zero"); // src/feLinearSystemMklPardiso.cpp:967-971
    _stateFresh = false;
  }
  void setMatrixToZero() { ok(b200_set_to_zero(_sys, 2), "b200_set_to_zero"); }
  void setResidualToZero() { ok(b200_set_to_zero(_sys, 1), "b200_set_to_zero"); }

  void assemble(const feSolution *sol, const bool assembleOnlyTransientMatrices = false)
  {
    upload(sol);
    ok(b200_assemble(_sys, _recomputeMatrix ? 3 : 1, assembleOnlyTransientMatrices), "b200_assemble");
  }
  void assembleMatrices(const feSolution *sol, const bool assembleOnlyTransientMatrices = false)
  {
    if(!_stateFresh || _opt.alwaysUpload) upload(sol);
    ok(b200_assemble(_sys, 2, assembleOnlyTransientMatrices), "b200_assemble");
  }
  void assembleResiduals(const feSolution *sol)
  {
    upload(sol);
    _stateFresh = true; // solveNewtonRaphson assembles the matrix from the same state (src/feNonLinearSolver.cpp:80-91)
    ok(b200_assemble(_sys, 1, 0), "b200_assemble");
  }

  void constrainEssentialComponents(const feSolution *sol)
  {
    if(!_constraintInit) initConstraints(sol);
    ok(b200_constrain(_sys), "b200_constrain");
  }
  void applyPeriodicity()
  {
    ok(b200_apply_periodicity(_sys), "b200_apply_periodicity");
  }
  void permute() {}

  bool solve(double *normDx, double *normResidual, double *normAxb, int *nIter)
  {
    b200_solver_options o{_rel_tol, _abs_tol, _div_tol, _max_iter, _opt.restart, _opt.preconditioner};
    const int           rc = b200_solve(_sys, &o, &_lastInfo);
    *normDx       = _lastInfo.norm_dx;
    *normResidual = _lastInfo.norm_rhs;
    *normAxb      = _lastInfo.norm_axb;
    *nIter        = _lastInfo.iterations;
    _numSolves++;
    _totalKrylovIterations += _lastInfo.iterations;
    if(rc < 0) return fail("b200_solve");
    if(!_lastInfo.converged) {
      feWarning("feLinearSystemB200: GMRES did not reach the tolerance in %d iterations (relative residual %.3e)",
                _lastInfo.iterations, _lastInfo.rel_residual);
      return false;
    }
    return true;
  }

  void correctSolution(feSolution *sol, const bool correctSolutionDot = false)
  {
    std::vector<double> &v = correctSolutionDot ? sol->getSolutionDot() : sol->getSolution();
    ok(b200_correct_solution(_sys, v.data(), correctSolutionDot ? 1 : 0), "b200_correct_solution");
    _stateFresh = false;
  }

  void assignResidualToDCResidual(feSolutionContainer *solContainer)
  {
    std::vector<double> rhs(_nInc);
    if(ok(b200_get_rhs(_sys, rhs.data()), "b200_get_rhs"))
      for(feInt i = 0; i < _nInc; ++i) solContainer->_fResidual[0][i] = rhs[i];
  }
  void applyCorrectionToResidual(double coeff, std::vector<double> &d) { ok(b200_axpy_rhs(_sys, coeff, d.data()), "b200_axpy_rhs"); }

  void viewMatrix() const
  {
    int64_t n, nnz;
    b200_get_pattern_size(_sys, &n, &nnz);
    std::vector<int64_t> ia(n + 1);
    std::vector<int32_t> ja(nnz);
    std::vector<double>  v(nnz);
    b200_get_pattern(_sys, ia.data(), ja.data());
    b200_get_matrix_values(_sys, v.data());
    for(int64_t i = 0; i < n; ++i)
      for(int64_t k = ia[i]; k < ia[i + 1]; ++k) printf("A(%ld, %d) = %+-10.10e\n", (long)i, ja[k], v[k]);
  }
  void viewRHS() const
  {
    std::vector<double> r(_nInc);
    b200_get_rhs(_sys, r.data());
    for(feInt i = 0; i < _nInc; ++i) printf("b(%ld) = %+-10.10e\n", (long)i, r[i]);
  }
  void viewResidual() const
  {
    std::vector<double> r(_nInc);
    b200_get_du(_sys, r.data());
    for(feInt i = 0; i < _nInc; ++i) printf("du(%ld) = %+-10.10e\n", (long)i, r[i]);
  }
  // one value per line in CSR order, "%+-10.20e" as src/feLinearSystemMklPardiso.cpp:455-485
  void writeMatrix(const std::string fileName, const double = 0.)
  {
    int64_t n, nnz;
    b200_get_pattern_size(_sys, &n, &nnz);
    std::vector<double> v(nnz);
    b200_get_matrix_values(_sys, v.data());
    FILE *f = fopen(fileName.c_str(), "w");
    if(!f) return;
    for(double x : v) fprintf(f, "%+-10.20e\n", x);
    fclose(f);
  }
  void writeRHS(const std::string fileName, const double = 0.)
  {
    std::vector<double> r(_nInc);
    b200_get_rhs(_sys, r.data());
    FILE *f = fopen(fileName.c_str(), "w");
    if(!f) return;
    for(double x : r) fprintf(f, "%+-10.20e\n", x);
    fclose(f);
  }
  void writeResidual(const std::string fileName, const double = 0.)
  {
    std::vector<double> r(_nInc);
    b200_get_du(_sys, r.data());
    FILE *f = fopen(fileName.c_str(), "w");
    if(!f) return;
    for(double x : r) fprintf(f, "%+-10.20e\n", x);
    fclose(f);
  }
};

// Factory in the style of createLinearSystem (src/feLinearSystem.cpp:36-64); the reference's own factory enum lives in
// the read-only tree, so the B200 backend ships its own entry point.
inline feStatus createLinearSystemB200(feLinearSystem *&system, const std::vector<feBilinearForm *> bilinearForms,
                                       const feMetaNumber *numbering, const feB200Options &opt = feB200Options())
{
  if(bilinearForms.empty()) return feErrorMsg(FE_STATUS_ERROR, "Cannot create a linear system without weak forms.");
  feLinearSystemB200 *s = new feLinearSystemB200(bilinearForms, numbering, opt);
  if(s->getStatus() != FE_STATUS_OK) {
    delete s;
    system = nullptr;
    return FE_STATUS_ERROR;
  }
  system = s;
  return FE_STATUS_OK;
}

#endif
