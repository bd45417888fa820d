// Compiled stand-in for `program mhd` (src_compressible/mhd.f90:43-293) over the C ABI of include/laps_b200.h:
// what the Fortran driver does around the hot path once its FFTW / MPI-transpose calls are replaced by the library
// (INTEGRATION.md), written in C++ because this image has no Fortran compiler.  One rank; the four source trees share the
// namelist syntax and are chosen with --tree compressible | compressible2d | incompressible | incompressible2d (the 2D trees:
// nz = 1, vardt every 20 steps, checkNan every 200, two-grid grid.dat, 2D/mhd.f90:22-25,237-252, 2D/mhdoutput.f90:45-63):
//   namelists (mhd.f90:30-53)  ->  laps_create            initial data (ifield = 3, ipert = 0 / 1; or a restart file)
//   Principal loop (mhd.f90:169-287): output at dtout / dtrms cadence, evolve, time += dt, evolve_radius, vardt
//   files in the reference's formats: grid.dat, parallel_info.dat (mhdoutput.f90:51-69), outNNN.dat (:72-131),
//   rms.dat (mhdrms.f90:25,48), EBM_info.dat (AEBmod.f90:75-85), log (mhd.f90:431-457)
// Every number comes from the library through plain pointers; nothing here computes physics.  The Python stand-in
// laps_b200/driver.py does the same (plus the other trees and several ranks); tests/test_cpp_driver.py holds the two to
// the same files.
//
//   g++ -std=c++17 -O2 -Iinclude integration/mhd_main.cpp -Llaps_b200/_lib -l:liblaps_b200.so -Wl,-rpath,$PWD/laps_b200/_lib -o mhd_main
//   ./mhd_main --input mhd.input --outdir run1 [--tree compressible2d] [--max-steps N]
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "laps_b200.h"

namespace {

// ---- namelists ("&group ... /", case-insensitive names, "!" comments; T / F / .true. / .false.; 1d-2) -----------------
typedef std::map<std::string, std::map<std::string, std::string>> Namelists;

std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }

Namelists read_namelists(const std::string& path) {
  Namelists nl;
  std::ifstream f(path);
  if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
  std::string line, cur;
  while (std::getline(f, line)) {
    line = line.substr(0, line.find('!'));
    std::string spaced;
    for (char c : line) {
      if (c == ',') spaced += ' ';
      else if (c == '=') spaced += " = ";
      else spaced += c;
    }
    std::istringstream ss(spaced);
    std::vector<std::string> tok;
    for (std::string t; ss >> t;) tok.push_back(t);
    for (size_t i = 0; i < tok.size(); ++i) {
      if (tok[i][0] == '&' && lower(tok[i]) != "&end") { cur = lower(tok[i].substr(1)); nl[cur]; continue; }
      if (tok[i] == "/" || lower(tok[i]) == "&end" || lower(tok[i]) == "$end") { cur.clear(); continue; }
      if (!cur.empty() && i + 2 < tok.size() && tok[i + 1] == "=") { nl[cur][lower(tok[i])] = tok[i + 2]; i += 2; }
    }
  }
  return nl;
}

struct Input {
  const Namelists& nl;
  const std::string* find(const char* group, const char* key) const {
    auto g = nl.find(lower(group));
    if (g == nl.end()) return nullptr;
    auto k = g->second.find(lower(key));
    return k == g->second.end() ? nullptr : &k->second;
  }
  double real(const char* group, const char* key, double dflt) const {
    const std::string* v = find(group, key);
    if (!v) return dflt;
    std::string s = *v;
    for (auto& c : s) if (c == 'd' || c == 'D') c = 'e';
    return std::strtod(s.c_str(), nullptr);
  }
  int integer(const char* group, const char* key, int dflt) const { return (int)std::lround(real(group, key, dflt)); }
  int logical(const char* group, const char* key, bool dflt) const {
    const std::string* v = find(group, key);
    if (!v) return dflt;
    const std::string s = lower(*v);
    return s == "t" || s == ".true." || s == ".t." || s == "true";
  }
};

// ---- Fortran-formatted output -------------------------------------------------------------------------------------------
std::string fmt_1pe16_8(double x) {   // "1pe16.8": E+dd, and +ddd without the E for three-digit exponents
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%.8E", x);
  std::string s(buf);
  const size_t e = s.find('E');
  const int ex = std::atoi(s.c_str() + e + 1);
  if (std::abs(ex) >= 100) { std::snprintf(buf, sizeof(buf), "%s%+04d", s.substr(0, e).c_str(), ex); s = buf; }
  return std::string(s.size() < 16 ? 16 - s.size() : 0, ' ') + s;
}

void put_record(std::FILE* f, const void* p, int32_t bytes) {   // one sequential unformatted record
  std::fwrite(&bytes, 4, 1, f); std::fwrite(p, 1, (size_t)bytes, f); std::fwrite(&bytes, 4, 1, f);
}

std::string out_name(int iout) { char b[32]; std::snprintf(b, sizeof(b), "out%03d.dat", iout); return b; }

struct Run {
  laps_handle h = nullptr;
  laps_params p{};
  std::string outdir;
  int nx = 0, ny = 0, nz = 0;
  double time = 0.0, radius = 0.0, Ur = 0.0;
  int output_primitive = 1;
  long istep = 0;
  std::chrono::steady_clock::time_point clock0 = std::chrono::steady_clock::now();
  std::vector<double> buf;   // 8 fields, the layout of uu(ix,iy,iz,ivar)

  std::string path(const std::string& name) const { return outdir + "/" + name; }
  double clock() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - clock0).count(); }
  void ck(int rc, const char* what) const {
    if (rc != 0) { std::fprintf(stderr, "%s failed: %s\n", what, laps_last_error(h)); std::exit(1); }
  }
  void output_uu(int iout) {   // mhdoutput.f90:72-131 on one rank: record(real(time,4)), then uu(nx,ny,nz,nvar) as float64
    ck(laps_get_output(h, buf.data(), output_primitive), "laps_get_output");
    std::FILE* f = std::fopen(path(out_name(iout)).c_str(), "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", path(out_name(iout)).c_str()); std::exit(2); }
    const float t4 = (float)time;
    put_record(f, &t4, 4);
    std::fwrite(buf.data(), 8, buf.size(), f);
    std::fclose(f);
  }
  void output_rms() {          // mhdrms.f90:39-51, format (f12.6,2x,19(1pe16.8))
    double r[19];
    ck(laps_rms(h, r), "laps_rms");
    std::FILE* f = std::fopen(path("rms.dat").c_str(), "a");
    std::fprintf(f, "%12.6f  ", time);
    for (double v : r) std::fputs(fmt_1pe16_8(v).c_str(), f);
    std::fputc('\n', f);
    std::fclose(f);
  }
  void output_aeb() {          // AEBmod.f90:75-85, format (3(1pe16.8))
    std::FILE* f = std::fopen(path("EBM_info.dat").c_str(), "a");
    std::fprintf(f, "%s%s%s\n", fmt_1pe16_8(time).c_str(), fmt_1pe16_8(radius).c_str(), fmt_1pe16_8(Ur).c_str());
    std::fclose(f);
  }
  void write_log(double dt) {  // mhd.f90:431-457
    const double c = clock();
    const double hour = std::floor(c / 3600.0), minute = std::floor((c / 3600.0 - hour) * 60);
    const double second = std::floor(((c / 3600.0 - hour) * 60 - minute) * 60);
    std::FILE* f = std::fopen(path("log").c_str(), "w");
    std::fprintf(f, "   Simulation time:%8.4f\n dt:  %.17g\n   Real time (sec):%15.2f\n", time, dt, c);
    std::fprintf(f, "   Real time (hh,mm,ss):%3dh%3dm%3ds\n   Iterations     :%8ld\n tasks: %12d\n", (int)hour, (int)minute, (int)second, istep, 1);
    std::fclose(f);
  }
};

}  // namespace

int main(int argc, char** argv) {
  std::string input = "mhd.input", outdir = ".", tree = "compressible";
  long max_steps = -1;
  bool echo = true;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--input" && i + 1 < argc) input = argv[++i];
    else if (a == "--outdir" && i + 1 < argc) outdir = argv[++i];
    else if (a == "--tree" && i + 1 < argc) tree = argv[++i];
    else if (a == "--max-steps" && i + 1 < argc) max_steps = std::atol(argv[++i]);
    else if (a == "--quiet") echo = false;
    else { std::fprintf(stderr, "usage: mhd_main [--input mhd.input] [--outdir DIR] [--tree compressible|compressible2d|incompressible|incompressible2d] [--max-steps N] [--quiet]\n"); return 2; }
  }
  if (tree != "compressible" && tree != "compressible2d" && tree != "incompressible" && tree != "incompressible2d") {
    std::fprintf(stderr, "unknown --tree %s\n", tree.c_str()); return 2;
  }
  const bool two_d = tree.size() > 2 && tree.compare(tree.size() - 2, 2, "2d") == 0;
  const bool incompressible = tree.compare(0, 6, "incomp") == 0;
  const Namelists nl = read_namelists(input);
  const Input in{nl};
  Run r;
  r.outdir = outdir;

  // ---- namelists -> laps_params (module defaults where a key is absent: mhdinit.f90:5-54, dealiasing.f90:9-10, AEBmod.f90:10-12)
  laps_params& p = r.p;
  p.abi_version = LAPS_ABI_VERSION;
  p.nx = in.integer("grid", "nx", 128); p.ny = in.integer("grid", "ny", 128); p.nz = in.integer("grid", "nz", 64);
  p.Lx = in.real("grid", "Lx", 1.0); p.Ly = in.real("grid", "Ly", 1.0); p.Lz = in.real("grid", "Lz", 1.0);
  p.adiabatic_index = in.real("phys", "adiabatic_index", 5.0 / 3.0);
  p.if_resis = in.logical("phys", "if_resis", false); p.resistivity = in.real("phys", "resistivity", 0.0);
  p.if_visc = in.logical("phys", "if_visc", false); p.viscosity = in.real("phys", "viscosity", 0.0);
  p.if_resis_exp = in.logical("numerical", "if_resis_exp", false); p.if_visc_exp = in.logical("numerical", "if_visc_exp", false);
  p.if_conserve_background = in.logical("numerical", "if_conserve_background", false);
  p.cfl = in.real("numerical", "cfl", 0.5); p.dealias_option = in.integer("numerical", "dealias_option", 2);
  p.afx = in.real("numerical", "afx", 0.495); p.afy = in.real("numerical", "afy", 0.495); p.afz = in.real("numerical", "afz", 0.495);
  p.if_AEB = in.logical("AEB", "if_AEB", false); p.if_corotating = in.logical("AEB", "if_corotating", false);
  p.radius0 = in.real("AEB", "radius0", 30.0); p.Ur0 = in.real("AEB", "Ur0", 0.0);
  p.corotating_angle = in.real("AEB", "corotating_angle", 0.0);
  p.if_hall = in.logical("Hall", "if_Hall", false); p.ion_inertial_length = in.real("Hall", "ion_inertial_length", 0.0);
  p.rank = 0; p.nranks = 1; p.device = 0; p.ndim = 3; p.rho0 = 1.0;
  if (two_d) {                                        // 2D/mhd.f90:23,43,44
    p.ndim = 2; p.nz = 1;
    p.if_limit_dt_increase = in.logical("numerical", "if_limit_dt_increase", false);
    if (!incompressible) {
      p.if_z_radial = in.logical("AEB", "if_z_radial", false);
      p.if_external_force = in.logical("pert", "if_external_force", false);
    }
  }
  if (incompressible) p.incompressible = 1;           // rho0 = 1 (src_incompressible/mhdinit.f90:15)
  const long dstep_calcdt = two_d ? 20 : 1;           // 2D/mhd.f90:22,237-240
  const long dstep_checknan = two_d ? 200 : 0;        // 2D/mhd.f90:25,242-252
  const double tmax = in.real("genr", "tmax", 1.0), dtout = in.real("genr", "dtout", 1.0), dtrms = in.real("genr", "dtrms", 1.0);
  r.output_primitive = in.logical("genr", "output_primitive", true);
  const bool if_restart = in.logical("genr", "if_restart", false);
  const int n_start = in.integer("genr", "n_start", 0);
  const long dstep_checksave = 40;                    // mhd.f90:20
  const double delta_clocktime_output = 60.0 * 50.0;  // mhd.f90:16

  if (laps_create(&p, &r.h) != 0) { std::fprintf(stderr, "laps_create: %s\n", laps_last_error(nullptr)); return 1; }
  r.nx = p.nx; r.ny = p.ny; r.nz = p.nz;
  const size_t npts = (size_t)p.nx * p.ny * p.nz;
  r.buf.assign(8 * npts, 0.0);
  r.Ur = p.if_AEB ? p.Ur0 : 0.0;                      // mhd.f90:88-90
  r.radius = p.radius0;

  // ---- initial data (mhd.f90:101-122) ------------------------------------------------------------------------------------
  std::vector<double>& prim = r.buf;                  // rho, ux, uy, uz, bx, by, bz, p
  if (if_restart) {                                   // restart.f90:17-63: the outNNN.dat of an output_primitive run
    std::FILE* f = std::fopen(r.path(out_name(n_start)).c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", r.path(out_name(n_start)).c_str()); return 2; }
    int32_t len = 0; float t4 = 0.f;
    if (std::fread(&len, 4, 1, f) != 1 || std::fread(&t4, 4, 1, f) != 1 || std::fread(&len, 4, 1, f) != 1 ||
        std::fread(prim.data(), 8, prim.size(), f) != prim.size()) { std::fprintf(stderr, "short restart file\n"); return 2; }
    std::fclose(f);
    r.time = (double)t4;
    r.ck(laps_set_time(r.h, r.time), "laps_set_time");   // mhd.f90:101-103
    r.radius = p.radius0 + r.Ur * r.time;
  } else {
    const int ifield = in.integer("field", "ifield", 3), ipert = in.integer("pert", "ipert", 0);
    if (ifield != 3 || (ipert != 0 && ipert != 1)) {
      std::fprintf(stderr, "built-in initial data: ifield = 3 with ipert = 0 or 1 (anything else: restart from an outNNN.dat)\n"); return 2;
    }
    const double b0[3] = {in.real("field", "Bx0", 0.0), in.real("field", "By0", 0.0), in.real("field", "Bz0", 0.0)};
    const double press0 = in.real("field", "press0", 1.0);
    for (size_t i = 0; i < npts; ++i) {               // background_fields_initialize case 3 (mhdinit.f90:183-260)
      prim[i] = 1.0;
      prim[4 * npts + i] = b0[0]; prim[5 * npts + i] = b0[1]; prim[6 * npts + i] = b0[2];
      prim[7 * npts + i] = press0;
    }
    if (ipert == 1) {                                 // circularly polarised Alfven wave along x (mhdinit.f90:328-342)
      const double pi = 3.141592653589793, db0 = in.real("pert", "db0", 0.1);
      const double kx = 2 * pi / p.Lx * in.integer("pert", "wave_number_jet", 1);
      const double ang = p.if_corotating ? p.corotating_angle : 0.0, ca = std::cos(ang), sa = std::sin(ang);
      for (size_t i = 0; i < npts; ++i) {
        const double x = (double)(i % p.nx) * (p.Lx / p.nx), s = std::sin(kx * x), c = std::cos(kx * x);
        const double rs = std::sqrt(prim[i]);
        prim[6 * npts + i] -= db0 * s;          prim[3 * npts + i] += db0 / rs * s;
        prim[1 * npts + i] += db0 / rs * c * sa; prim[4 * npts + i] -= db0 * c * sa;
        prim[2 * npts + i] += db0 / rs * c * ca; prim[5 * npts + i] -= db0 * c * ca;
      }
    }
  }
  r.ck(laps_set_primitive(r.h, prim.data()), "laps_set_primitive");   // mhd.f90:121-122
  double dt = 0.0;
  r.ck(laps_vardt(r.h, &dt), "laps_vardt");                           // :135-136

  // ---- grid.dat, parallel_info.dat (mhdoutput.f90:51-69); rms.dat / EBM_info.dat are opened for append -----------------
  {
    std::vector<float> g;                             // the 2D trees write nx, ny and two grids; npe, nvar (2D/mhdoutput.f90:45-63)
    const float dims[3] = {(float)p.nx, (float)p.ny, (float)p.nz};
    for (int i = 0; i < p.nx; ++i) g.push_back((float)(i * (p.Lx / p.nx)));
    for (int i = 0; i < p.ny; ++i) g.push_back((float)(i * (p.Ly / p.ny)));
    if (!two_d) for (int i = 0; i < p.nz; ++i) g.push_back((float)(i * (p.Lz / p.nz)));
    std::FILE* f = std::fopen(r.path("grid.dat").c_str(), "wb");
    if (!f) { std::fprintf(stderr, "cannot write into %s\n", outdir.c_str()); return 2; }
    put_record(f, dims, two_d ? 8 : 12); put_record(f, g.data(), (int32_t)(4 * g.size()));
    std::fclose(f);
    const float info3[4] = {1.f, 1.f, 1.f, 8.f}, info2[2] = {1.f, 8.f};   // npe, iproc, jproc, nvar / npe, nvar
    f = std::fopen(r.path("parallel_info.dat").c_str(), "wb");
    if (two_d) put_record(f, info2, 8); else put_record(f, info3, 16);
    std::fclose(f);
    std::fclose(std::fopen(r.path("rms.dat").c_str(), "a"));
    std::fclose(std::fopen(r.path("EBM_info.dat").c_str(), "a"));
  }

  // ---- mhd.f90:142-166 -----------------------------------------------------------------------------------------------------
  const double dtlog = std::fmin(dtout, dtrms) / 10.0;
  int iout = n_start;
  double tout = r.time + dtout, toutrms = r.time + dtrms, tlog = r.time + dtlog;
  double max_divb = 0.0;
  auto rms_block = [&]() {
    r.ck(laps_max_divb(r.h, &max_divb), "laps_max_divb");
    if (echo) std::printf("      OUTPUT RMS at time:  %10.4f, max(div B) = %10.2E, dt = %12.4E\n", r.time, max_divb, dt);
    r.output_rms();
    r.output_aeb();
  };
  r.output_uu(iout++);
  rms_block();
  double clocktime_output = delta_clocktime_output;
  std::vector<double> force;

  // ---- Principal (mhd.f90:169-287) ---------------------------------------------------------------------------------------
  for (;;) {
    if (r.time >= tout) { r.output_uu(iout++); tout += dtout; }
    if (r.time >= toutrms) { rms_block(); toutrms += dtrms; }
    if (r.istep > 0 && r.istep % dstep_checksave == 0 && r.clock() >= clocktime_output) {   // wall-clock backup dump, :194-214
      if (echo) std::printf("   OUTPUT for backup at real time (sec):%15.2f  , time =   %10.4f\n", r.clock(), r.time);
      r.output_uu(999);
      clocktime_output += delta_clocktime_output;
    }
    if (dt < 1e-8) { r.output_uu(iout); r.output_rms(); r.output_aeb(); break; }          // :205-228
    if (p.if_external_force) {   // the user routine calc_external_force_real as shipped (2D/mhdrhs.f90:480-531): a Gaussian forcing of
      // B_z centred at x = Lx/2 whose y centre moves at speed 0.3, with its two periodic images; called with the step's time
      const double dBdt = 0.2, xc = 0.5 * p.Lx, w = 0.05 * p.Ly, yc = std::fmod(0.2 * p.Ly + 0.3 * r.time, p.Ly);
      force.resize((size_t)p.nx * p.ny);
      for (int iy = 0; iy < p.ny; ++iy)
        for (int ix = 0; ix < p.nx; ++ix) {
          const double x = ix * (p.Lx / p.nx), y = iy * (p.Ly / p.ny), fx = std::exp(-((x - xc) / w) * ((x - xc) / w));
          double v = dBdt * fx * std::exp(-((y - yc) / w) * ((y - yc) / w));
          v = v + dBdt * fx * std::exp(-((y - (yc + p.Ly)) / w) * ((y - (yc + p.Ly)) / w));
          v = v + dBdt * fx * std::exp(-((y - (yc - p.Ly)) / w) * ((y - (yc - p.Ly)) / w));
          force[(size_t)iy * p.nx + ix] = v;
        }
      r.ck(laps_set_external_force(r.h, force.data()), "laps_set_external_force");
    }
    r.ck(laps_evolve(r.h), "laps_evolve");                                                // :245
    r.time = r.time + dt;                                                                 // :246
    ++r.istep;
    r.ck(laps_set_time(r.h, r.time), "laps_set_time");                                    // :248 evolve_radius
    r.radius = p.radius0 + r.Ur * r.time;
    if (r.time >= tmax || (max_steps >= 0 && r.istep >= max_steps)) {                     // :250-276
      r.output_uu(iout);
      r.ck(laps_max_divb(r.h, &max_divb), "laps_max_divb");
      r.output_rms();
      r.output_aeb();
      break;
    }
    if (r.time >= tlog) { r.write_log(dt); tlog += dtlog; }
    if (r.istep % dstep_calcdt == 0) r.ck(laps_vardt(r.h, &dt), "laps_vardt");            // :285 (2D trees: every 20 steps)
    if (dstep_checknan && r.istep % dstep_checknan == 0) {                                // 2D/mhd.f90:242-252
      int32_t is_nan = 0;
      r.ck(laps_check_nan(r.h, &is_nan), "laps_check_nan");
      if (is_nan) { if (echo) std::printf(" NaN encountered!!! Exit the program at t = %10.4f\n", r.time); break; }
    }
  }
  r.write_log(dt);
  r.ck(laps_sync(r.h), "laps_sync");
  laps_destroy(r.h);                                                                      // :291-293
  return 0;
}
