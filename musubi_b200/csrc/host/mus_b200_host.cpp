// mus_b200_host.cpp -- a compiled host above the C ABI (include/musb200.h), no Python involved:
// what musubi.f90 / mus_program_module.fpp do around the hot path for a single-level run on one
// rank -- mus_initialize (mesh, level descriptor, boundaries, initial state), the time loop
// control%do_computation (mus_control_module.f90:507-701) with check_flow_status
// (mus_aux_module.f90:115-207: total density, NaN) at an interval, the restart dump
// (mus_restart_module.f90:57-166: the raw element-major doubles of the *.lsb file) and the
// performance report of mus_perf_measure (mus_tools_module.f90:474-560: MLUPS).
//
// The Fortran host binds the same entry points through musubi_b200/fortran/mus_b200_module.f90;
// this program exists because no Fortran compiler is available where the library is built, and
// it is what `make host` links to prove that the header and the exported symbols agree from a
// compiled language.  The initial state comes from a file (--state-in: nElems * QQ doubles,
// AOS, as mus_pdf_serialize orders them) or is the rest state rho = 1, u = 0.
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "musb200.h"
#include "treelm_box.h"

namespace {

[[noreturn]] void abortWith(const char *where, int rc) {   // tem_abort
  char msg[512] = "";
  musb200_last_error(msg, (int)sizeof msg);
  std::fprintf(stderr, "mus_b200_host: %s failed (code %d): %s\n", where, rc, msg);
  std::exit(rc == 0 ? 1 : rc);
}
#define CHK(call) do { const int rc_ = (call); if (rc_ != 0) abortWith(#call, rc_); } while (0)

struct Options {
  int level = 5, steps = 100, check = 0, device = 0;
  std::string layout = "d3q19", relaxation = "bgk", kind = "fluid", mesh = "periodic", variant = "standard";
  double omega = 1.7, lambda = 0.25, omegaBulk = -1.0, lid[3] = {0.05, 0.0, 0.0}, rhoOut = 1.0;
  std::string stateIn, stateOut, outlet = "pressure_expol";
};

void usage() {
  std::puts(
      "usage: mus_b200_host [--level L] [--mesh periodic|cavity|channel] [--layout d3q19|d3q27]\n"
      "                     [--relaxation bgk|trt|mrt] [--kind fluid|fluid_incompressible]\n"
      "                     [--omega w] [--lambda l] [--omega-bulk w] [--steps N] [--check-interval K]\n"
      "                     [--lid ux uy uz] [--outlet pressure_expol|pressure_antibounceback]\n"
      "                     [--state-in file] [--state-out file] [--device d]\n"
      "single-level, single-rank run of the per-level LBM time step through libmusb200.so");
}

bool parse(int argc, char **argv, Options &o) {
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&](const char *what) -> const char * {
      if (i + 1 >= argc) { std::fprintf(stderr, "mus_b200_host: %s needs a value\n", what); std::exit(2); }
      return argv[++i];
    };
    if (a == "--help" || a == "-h") return false;
    else if (a == "--level") o.level = std::atoi(next("--level"));
    else if (a == "--steps") o.steps = std::atoi(next("--steps"));
    else if (a == "--check-interval") o.check = std::atoi(next("--check-interval"));
    else if (a == "--device") o.device = std::atoi(next("--device"));
    else if (a == "--mesh") o.mesh = next("--mesh");
    else if (a == "--layout") o.layout = next("--layout");
    else if (a == "--relaxation") o.relaxation = next("--relaxation");
    else if (a == "--variant") o.variant = next("--variant");
    else if (a == "--kind") o.kind = next("--kind");
    else if (a == "--omega") o.omega = std::atof(next("--omega"));
    else if (a == "--lambda") o.lambda = std::atof(next("--lambda"));
    else if (a == "--omega-bulk") o.omegaBulk = std::atof(next("--omega-bulk"));
    else if (a == "--rho-out") o.rhoOut = std::atof(next("--rho-out"));
    else if (a == "--outlet") o.outlet = next("--outlet");
    else if (a == "--state-in") o.stateIn = next("--state-in");
    else if (a == "--state-out") o.stateOut = next("--state-out");
    else if (a == "--lid") { for (double &v : o.lid) v = std::atof(next("--lid")); }
    else { std::fprintf(stderr, "mus_b200_host: unknown option %s\n", a.c_str()); std::exit(2); }
  }
  return true;
}

// stencil weights in the reference's direction order, rest last (mus_scheme_layout_module.f90:699-705)
std::vector<double> weights(int QQ) {
  std::vector<double> w((size_t)QQ);
  if (QQ == 19) {
    for (int d = 0; d < 6; ++d) w[d] = 1.0 / 18.0;
    for (int d = 6; d < 18; ++d) w[d] = 1.0 / 36.0;
    w[18] = 1.0 / 3.0;
  } else {
    for (int d = 0; d < 6; ++d) w[d] = 2.0 / 27.0;
    for (int d = 6; d < 18; ++d) w[d] = 1.0 / 54.0;
    for (int d = 18; d < 26; ++d) w[d] = 1.0 / 216.0;
    w[26] = 8.0 / 27.0;
  }
  return w;
}

}  // namespace

int main(int argc, char **argv) {
  Options o;
  if (!parse(argc, argv, o)) { usage(); return 0; }
  const int meshKind = o.mesh == "periodic" ? 0 : o.mesh == "cavity" ? 1 : o.mesh == "channel" ? 2 : -1;
  if (meshKind < 0) { std::fprintf(stderr, "mus_b200_host: unknown mesh %s\n", o.mesh.c_str()); return 2; }

  // ---- mus_init_advRel_*: identify -> kernel --------------------------------------------
  int relax = 0, kind = 0, QQ = 0;
  CHK(musb200_scheme_select(o.kind.c_str(), o.relaxation.c_str(), o.variant.c_str(), o.layout.c_str(),
                            &relax, &kind, &QQ));
  // ---- one rank = one GPU ------------------------------------------------------------------
  CHK(musb200_init(0, 1, o.device, nullptr));

  // ---- treelm + mus_construct: the level descriptor ----------------------------------------
  void *mesh = musb200_mesh_box_create(o.level, QQ, meshKind, 0, 1, 1, 8);
  if (!mesh) { std::fprintf(stderr, "mus_b200_host: bad mesh parameters\n"); return 2; }
  int64_t info[16] = {0};
  musb200_mesh_info(mesh, info);
  const int nFluid = (int)info[0], nHalo = (int)info[1], nElems = (int)info[2], nSize = (int)info[3];
  const int nBcElems = (int)info[4], nBCs = (int)info[5];
  std::vector<int64_t> total((size_t)nElems), prop((size_t)nElems);
  std::vector<int32_t> neigh((size_t)QQ * nSize);
  musb200_mesh_total(mesh, total.data());
  musb200_mesh_property(mesh, prop.data());
  musb200_mesh_neigh(mesh, neigh.data());
  const int L = o.level;
  CHK(musb200_level_create(L, QQ, QQ, 4, nSize, nFluid, 0, 0, nHalo, neigh.data(), prop.data(), total.data()));
  CHK(musb200_set_relaxation(L, relax, kind, nullptr, o.omega, o.lambda, o.omegaBulk > 0.0 ? o.omegaBulk : o.omega));

  // ---- boundaries (mus_init_boundary) -------------------------------------------------------
  if (nBcElems > 0) {
    std::vector<int32_t> eb((size_t)nBcElems);
    musb200_mesh_bc_elembuffer(mesh, eb.data());
    CHK(musb200_bc_elembuffer(L, nBcElems, eb.data()));
  }
  for (int i = 0; i < nBCs; ++i) {
    int32_t sz[4];
    musb200_mesh_bc_info(mesh, i, sz);
    const int id = sz[0], genKind = sz[1], nBE = sz[2], nLinks = sz[3];
    std::vector<int32_t> elems((size_t)std::max(nBE, 1)), links((size_t)std::max(nLinks, 1)),
        outPos(links.size()), pib(links.size()), iDir(links.size());
    musb200_mesh_bc_lists(mesh, i, elems.data(), links.data(), outPos.data(), pib.data(), iDir.data());
    int bcKind = MUSB200_BC_WALL;
    if (genKind == 1) bcKind = MUSB200_BC_VELOCITY_BOUNCEBACK;
    if (genKind == 2) bcKind = o.outlet == "pressure_antibounceback" ? MUSB200_BC_PRESSURE_ANTIBOUNCEBACK
                                                                    : MUSB200_BC_PRESSURE_EXPOL;
    CHK(musb200_bc_register(L, id, bcKind, nLinks, links.data(), outPos.data(), pib.data(), iDir.data()));
    if (genKind == 2 && nBE > 0) {
      std::vector<int32_t> normalInd((size_t)nBE), pibe((size_t)nBE), neighPos((size_t)2 * nBE),
          ieol(links.size()), statePos(links.size());
      musb200_mesh_bc_elem_lists(mesh, i, normalInd.data(), pibe.data(), neighPos.data(), ieol.data(), statePos.data());
      CHK(musb200_bc_register_elems(L, id, nBE, elems.data(), pibe.data(), normalInd.data(), 2, neighPos.data(),
                                    ieol.data()));
      std::vector<double> rho((size_t)nBE, o.rhoOut);                     // pressure * cs2inv / fac%press
      CHK(musb200_bc_set_values(L, id, nBE, rho.data()));
    }
    if (genKind == 1 && nLinks > 0) {
      std::vector<double> vel((size_t)3 * nLinks);                          // constant st-fun, lattice units
      for (int l = 0; l < nLinks; ++l) for (int c = 0; c < 3; ++c) vel[(size_t)3 * l + c] = o.lid[c];
      CHK(musb200_bc_set_values(L, id, 3 * nLinks, vel.data()));
    }
  }

  // ---- initial condition (mus_init_pdf) or a dump to continue from --------------------------
  std::vector<double> state((size_t)nSize * QQ, 0.0);
  if (!o.stateIn.empty()) {
    FILE *f = std::fopen(o.stateIn.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "mus_b200_host: cannot open %s: %s\n", o.stateIn.c_str(), std::strerror(errno)); return 2; }
    const size_t want = (size_t)nElems * QQ, got = std::fread(state.data(), sizeof(double), want, f);
    std::fclose(f);
    if (got != want && got != (size_t)nFluid * QQ) {
      std::fprintf(stderr, "mus_b200_host: %s holds %zu doubles, the level needs %zu\n", o.stateIn.c_str(), got, want);
      return 2;
    }
  } else {
    const std::vector<double> w = weights(QQ);
    for (int e = 0; e < nElems; ++e)
      for (int d = 0; d < QQ; ++d) state[(size_t)e * QQ + d] = w[(size_t)d];   // f_eq(rho = 1, u = 0)
  }
  CHK(musb200_state_upload(L, 2, state.data()));
  CHK(musb200_set_now_next(L, 1, 2));
  CHK(musb200_state_copy_next_to_now(L));
  // mus_init_flow: auxField of the fluid elements from the state, halos / ghosts filled
  // (pressure_expol reads the auxField of the previous step in its very first call)
  CHK(musb200_fill_helper_elements(L, L));

  // ---- the time loop -------------------------------------------------------------------------
  double mass0 = 0.0, vmax = 0.0;
  int nan = 0;
  CHK(musb200_reduce(L, &mass0, &vmax, &nan));
  CHK(musb200_synchronize());
  const auto t0 = std::chrono::steady_clock::now();
  int done = 0;
  while (done < o.steps) {
    const int n = o.check > 0 ? std::min(o.check, o.steps - done) : o.steps - done;
    CHK(musb200_step(L, L, n));
    done += n;
    if (o.check > 0) {                                  // check_flow_status
      double m = 0.0;
      CHK(musb200_reduce(L, &m, &vmax, &nan));
      std::printf("iter %8d  total density %.15e  max |u| %.6e%s\n", done, m, vmax, nan ? "  NaN!" : "");
      if (nan) { std::fprintf(stderr, "mus_b200_host: NaN detected, aborting\n"); return 3; }
    }
  }
  CHK(musb200_synchronize());
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double mass = 0.0;
  CHK(musb200_reduce(L, &mass, &vmax, &nan));
  std::printf("level %d  %s %s %s  %d elements  %d steps\n", L, o.kind.c_str(), o.relaxation.c_str(),
              o.layout.c_str(), nFluid, o.steps);
  std::printf("total density %.15e -> %.15e (relative change %.3e)  max |u| %.6e  nan %d\n", mass0, mass,
              mass0 != 0.0 ? mass / mass0 - 1.0 : 0.0, vmax, nan);
  std::printf("MLUPS %.1f (wall clock incl. checks, %.3f s)\n",
              secs > 0.0 ? (double)nFluid * o.steps / secs / 1e6 : 0.0, secs);

  // ---- restart dump: the payload of <sim>_<stamp>.lsb ------------------------------------------
  if (!o.stateOut.empty()) {
    std::vector<int32_t> lp((size_t)nFluid);
    for (int e = 0; e < nFluid; ++e) lp[(size_t)e] = e + 1;   // single level: tree order = total list order
    std::vector<double> buf((size_t)nFluid * QQ);
    CHK(musb200_pdf_serialize(nFluid, total.data(), lp.data(), buf.data()));
    FILE *f = std::fopen(o.stateOut.c_str(), "wb");
    if (!f || std::fwrite(buf.data(), sizeof(double), buf.size(), f) != buf.size()) {
      std::fprintf(stderr, "mus_b200_host: cannot write %s\n", o.stateOut.c_str());
      return 2;
    }
    std::fclose(f);
  }
  CHK(musb200_level_destroy(L));
  musb200_mesh_destroy(mesh);
  CHK(musb200_finalize());
  return nan ? 3 : 0;
}
