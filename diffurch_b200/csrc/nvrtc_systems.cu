// User systems from SOURCE: `Solver::equation` takes any closure (solver.rs:135-140); its CUDA
// restatement can be handed over as text and is compiled at run time with NVRTC for sm_100a — no nvcc
// on the machine, and only the kernel instantiation a problem actually needs is built:
//
//   (user functor, exact tableau sparsity, arithmetic contract, trace policy, members per thread)
//       -> ode_fixed_kernel<...>            hot path #1, registers only
//   (user functor, exact tableau shape, locators yes/no, ODEHIST)
//       -> integrate_general_kernel<...>    every other Solver feature
//
// compiled from the SAME kernel templates (csrc/dev/*.cuh) as the built-in systems and the nvcc-built
// plugins, so results are bit-identical to those.  libnvrtc and the driver API are resolved with
// dlopen at first use; compiled CUBINs are cached in memory and on disk (DFB_NVRTC_CACHE, default
// $TMPDIR/diffurch_b200_nvrtc) under a hash of the source, the instantiation and the library build.
#define DFB_ARITH 1
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "dev/integrator.cuh"
#include "dev/ode_fixed.cuh"
#include "host/launch_geometry.hpp"
#include "launch_table.hpp"
#include "nvrtc_systems.hpp"

namespace {

// ---- libnvrtc / libcuda, resolved at run time ----------------------------------------------------
struct Api {
  void* nvrtc = nullptr;
  void* cuda = nullptr;
  decltype(&nvrtcCreateProgram) createProgram = nullptr;
  decltype(&nvrtcDestroyProgram) destroyProgram = nullptr;
  decltype(&nvrtcCompileProgram) compileProgram = nullptr;
  decltype(&nvrtcGetProgramLogSize) getLogSize = nullptr;
  decltype(&nvrtcGetProgramLog) getLog = nullptr;
  decltype(&nvrtcGetCUBINSize) getCubinSize = nullptr;
  decltype(&nvrtcGetCUBIN) getCubin = nullptr;
  decltype(&nvrtcGetPTXSize) getPtxSize = nullptr;
  decltype(&nvrtcGetPTX) getPtx = nullptr;
  decltype(&nvrtcAddNameExpression) addName = nullptr;
  decltype(&nvrtcGetLoweredName) loweredName = nullptr;
  decltype(&cuModuleLoadData) moduleLoadData = nullptr;
  decltype(&cuModuleGetFunction) moduleGetFunction = nullptr;
  decltype(&cuLaunchKernel) launchKernel = nullptr;
  decltype(&cuFuncSetAttribute) funcSetAttribute = nullptr;
  decltype(&cuGetErrorString) getErrorString = nullptr;
  std::string error;
};

template <class F>
bool sym(void* lib, const char* name, F* out) {
  *out = reinterpret_cast<F>(dlsym(lib, name));
  return *out != nullptr;
}

Api* api(bool need_driver) {
  static Api a;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!a.nvrtc) {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (const char* n : names)
      if ((a.nvrtc = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!a.nvrtc) {
      a.error = "libnvrtc.so.12 not found (dlopen)";
      return &a;
    }
    bool ok = sym(a.nvrtc, "nvrtcCreateProgram", &a.createProgram) && sym(a.nvrtc, "nvrtcDestroyProgram", &a.destroyProgram) &&
              sym(a.nvrtc, "nvrtcCompileProgram", &a.compileProgram) && sym(a.nvrtc, "nvrtcGetProgramLogSize", &a.getLogSize) &&
              sym(a.nvrtc, "nvrtcGetProgramLog", &a.getLog) && sym(a.nvrtc, "nvrtcGetCUBINSize", &a.getCubinSize) &&
              sym(a.nvrtc, "nvrtcGetCUBIN", &a.getCubin) && sym(a.nvrtc, "nvrtcGetPTXSize", &a.getPtxSize) &&
              sym(a.nvrtc, "nvrtcGetPTX", &a.getPtx) && sym(a.nvrtc, "nvrtcAddNameExpression", &a.addName) &&
              sym(a.nvrtc, "nvrtcGetLoweredName", &a.loweredName);
    if (!ok) a.error = "libnvrtc: missing symbol";
  }
  if (need_driver && !a.cuda) {
    a.cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!a.cuda) {
      a.error = "libcuda.so.1 not found (no NVIDIA driver: diffurch_b200 has no CPU execution path)";
      return &a;
    }
    bool ok = sym(a.cuda, "cuModuleLoadData", &a.moduleLoadData) && sym(a.cuda, "cuModuleGetFunction", &a.moduleGetFunction) &&
              sym(a.cuda, "cuLaunchKernel", &a.launchKernel) && sym(a.cuda, "cuFuncSetAttribute", &a.funcSetAttribute) &&
              sym(a.cuda, "cuGetErrorString", &a.getErrorString);
    if (!ok) a.error = "libcuda: missing symbol";
  }
  return &a;
}

// ---- registry of source systems ---------------------------------------------------------------
struct SourceSys {
  int system_id;
  std::string source, struct_name;
  dfb_nvrtc_sysinfo info;
};
std::vector<SourceSys*>& sources() {
  static std::vector<SourceSys*> v;
  return v;
}
const SourceSys* find_source(int system_id) {
  for (const SourceSys* s : sources())
    if (s->system_id == system_id) return s;
  return nullptr;
}

struct Stats {
  unsigned compiled = 0, cache_hits = 0;
  double total_ms = 0.0, last_ms = 0.0;
} g_stats;
std::string g_error;

std::string csrc_dir() {
  if (const char* e = std::getenv("DFB_CSRC_DIR")) return e;
  Dl_info info;
  if (dladdr(reinterpret_cast<void*>(&csrc_dir), &info) && info.dli_fname) {
    std::string p = info.dli_fname;  // .../diffurch_b200/libdiffurch_b200.so
    const size_t slash = p.rfind('/');
    return (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/csrc";
  }
  return "diffurch_b200/csrc";
}

unsigned long long fnv1a(const std::string& s, unsigned long long h = 1469598103934665603ull) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}

std::string cache_dir() {
  std::string d;
  if (const char* e = std::getenv("DFB_NVRTC_CACHE")) d = e;
  else {
    const char* t = std::getenv("TMPDIR");
    d = std::string(t ? t : "/tmp") + "/diffurch_b200_nvrtc";
  }
  mkdir(d.c_str(), 0755);
  return d;
}

std::string library_stamp() {  // a rebuilt library (changed kernel templates) must not reuse old CUBINs
  Dl_info info;
  struct stat st;
  if (dladdr(reinterpret_cast<void*>(&library_stamp), &info) && info.dli_fname && stat(info.dli_fname, &st) == 0)
    return std::to_string((long long)st.st_mtime) + ":" + std::to_string((long long)st.st_size);
  return "?";
}

// Compile `main_text` (+ the user's functor as the virtual header dfb_user_system.cuh).  kind 0: CUBIN for
// the kernel named by `name_expr` (its lowered name is returned); kind 1: PTX (registration probe).
bool compile(const SourceSys& S, int arith, const std::string& main_text, const std::string& name_expr, int kind,
             std::vector<char>* image, std::string* lowered) {
  Api* A = api(false);
  if (!A->error.empty()) {
    g_error = A->error;
    return false;
  }
  const std::string dir = csrc_dir();
  std::string user = S.source;
  if (user.find("user_system.cuh") == std::string::npos)
    user = "#include \"user_system.cuh\"\nnamespace DFB_NS {\n" + user + "\n}  // namespace DFB_NS\n";
  const std::string key_text = user + "\x1f" + main_text + "\x1f" + name_expr + "\x1f" + std::to_string(arith) + "\x1f" +
                               std::to_string(kind) + "\x1f" + library_stamp();
  char key[40];
  std::snprintf(key, sizeof(key), "%016llx", fnv1a(key_text));
  const std::string cache_file = cache_dir() + "/" + key + (kind == 0 ? ".cubin" : ".ptx");
  const std::string name_file = cache_dir() + "/" + key + ".name";
  if (!std::getenv("DFB_NVRTC_NO_DISK_CACHE")) {
    FILE* f = std::fopen(cache_file.c_str(), "rb");
    FILE* g = kind == 0 ? std::fopen(name_file.c_str(), "rb") : nullptr;
    if (f && (kind != 0 || g)) {
      std::fseek(f, 0, SEEK_END);
      const long n = std::ftell(f);
      std::fseek(f, 0, SEEK_SET);
      image->resize(size_t(n));
      const bool ok = std::fread(image->data(), 1, size_t(n), f) == size_t(n);
      std::fclose(f);
      if (g) {
        char buf[2048] = {0};
        const size_t m = std::fread(buf, 1, sizeof(buf) - 1, g);
        std::fclose(g);
        lowered->assign(buf, m);
      }
      if (ok) {
        ++g_stats.cache_hits;
        return true;
      }
    } else {
      if (f) std::fclose(f);
      if (g) std::fclose(g);
    }
  }
  const auto t0 = std::chrono::steady_clock::now();
  const char* header_names[] = {"dfb_user_system.cuh"};
  const char* header_texts[] = {user.c_str()};
  nvrtcProgram prog = nullptr;
  if (A->createProgram(&prog, main_text.c_str(), "dfb_nvrtc_unit.cu", 1, header_texts, header_names) != NVRTC_SUCCESS) {
    g_error = "nvrtcCreateProgram failed";
    return false;
  }
  if (!name_expr.empty()) A->addName(prog, name_expr.c_str());
  const std::string inc1 = "-I" + dir, inc2 = "-I" + dir + "/../../include";
  const std::string d_arith = "-DDFB_ARITH=" + std::to_string(arith);
  const std::string d_sys = "-DDFB_PLUGIN_SYSTEM=" + S.struct_name;
  std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "-lineinfo",
                                   inc1.c_str(), inc2.c_str(), "-I/usr/local/cuda/include", d_arith.c_str(),
                                   "-DDFB_PLUGIN_HEADER=\"dfb_user_system.cuh\"", d_sys.c_str()};
  if (arith == 0) opts.push_back("--fmad=false");
  const nvrtcResult rc = A->compileProgram(prog, int(opts.size()), opts.data());
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    A->getLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) A->getLog(prog, &log[0]);
    g_error = "NVRTC compilation failed:\n" + log;
    A->destroyProgram(&prog);
    return false;
  }
  size_t n = 0;
  if (kind == 0) {
    A->getCubinSize(prog, &n);
    image->resize(n);
    A->getCubin(prog, image->data());
    const char* low = nullptr;
    if (A->loweredName(prog, name_expr.c_str(), &low) != NVRTC_SUCCESS || !low) {
      g_error = "nvrtcGetLoweredName failed for " + name_expr;
      A->destroyProgram(&prog);
      return false;
    }
    *lowered = low;
  } else {
    A->getPtxSize(prog, &n);
    image->resize(n);
    A->getPtx(prog, image->data());
  }
  A->destroyProgram(&prog);
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  ++g_stats.compiled;
  g_stats.total_ms += ms;
  g_stats.last_ms = ms;
  if (!std::getenv("DFB_NVRTC_NO_DISK_CACHE")) {
    if (FILE* f = std::fopen(cache_file.c_str(), "wb")) {
      std::fwrite(image->data(), 1, image->size(), f);
      std::fclose(f);
    }
    if (kind == 0)
      if (FILE* g = std::fopen(name_file.c_str(), "wb")) {
        std::fwrite(lowered->data(), 1, lowered->size(), g);
        std::fclose(g);
      }
  }
  return true;
}

// ---- kernels: compile on demand, load once ---------------------------------------------------------
std::map<std::string, CUfunction>& loaded() {
  static std::map<std::string, CUfunction> m;
  return m;
}

const char* ns_of(int arith) { return arith == 0 ? "dfb_strict" : "dfb_fast"; }

// returns nullptr with g_error set on failure
CUfunction get_kernel(const SourceSys& S, int arith, const char* header, const std::string& name_expr) {
  const std::string key = std::to_string(S.system_id) + "|" + std::to_string(arith) + "|" + name_expr;
  auto it = loaded().find(key);
  if (it != loaded().end()) return it->second;
  std::vector<char> cubin;
  std::string lowered;
  const std::string main_text = std::string("#include \"") + header + "\"\n";
  if (!compile(S, arith, main_text, name_expr, 0, &cubin, &lowered)) return nullptr;
  Api* A = api(true);
  if (!A->error.empty()) {
    g_error = A->error;
    return nullptr;
  }
  cudaFree(nullptr);  // make the primary context current for the driver API
  CUmodule mod = nullptr;
  CUresult rc = A->moduleLoadData(&mod, cubin.data());
  CUfunction fn = nullptr;
  if (rc == CUDA_SUCCESS) rc = A->moduleGetFunction(&fn, mod, lowered.c_str());
  if (rc != CUDA_SUCCESS) {
    const char* msg = nullptr;
    A->getErrorString(rc, &msg);
    g_error = std::string("loading the NVRTC-built kernel failed: ") + (msg ? msg : "?");
    return nullptr;
  }
  loaded()[key] = fn;
  return fn;
}

std::string hex64(unsigned long long v) {
  char b[32];
  std::snprintf(b, sizeof(b), "0x%llxull", v);
  return b;
}

// ---- shims with the signatures of the launch table -----------------------------------------------
template <int ARITH>
int nvrtc_ode_fixed_shim(int system_id, const dfb_tableau& rk, double t0, double t1, double h, size_t n,
                         const double* y0, const double* params, double* t_final, double* y_final,
                         unsigned long long* steps, unsigned long long* rhs_evals, uint32_t* status, int trace_every,
                         int trace_capacity, double* trace_t, double* trace_y, uint32_t* n_samples,
                         cudaStream_t stream) {
  const SourceSys* S = find_source(system_id);
  if (!S || S->info.uses_history || S->info.reads_d) return -1;
  dfb_fast::OdeFixedArgs A;  // the same POD in both arithmetic namespaces
  std::memset(&A, 0, sizeof(A));
  unsigned long long amask = 0;
  unsigned bmask = 0;
  for (int i = 0; i < DFB_MAX_STAGES; ++i) {
    for (int j = 0; j < DFB_MAX_STAGES; ++j) {
      const double a = rk.a[i * DFB_MAX_STAGES + j];
      A.a[i * DFB_MAX_STAGES + j] = a;
      A.ha[i * DFB_MAX_STAGES + j] = h * a;
      if (i < rk.S && j < i && a != 0.0) amask |= dfb_fast::amask_bit(i, j);
    }
    A.b[i] = rk.b[i];
    A.c[i] = rk.c[i];
    A.hb[i] = h * rk.b[i];
    A.hc[i] = rk.c[i] * h;
    if (i < rk.S && rk.b[i] != 0.0) bmask |= 1u << i;
  }
  A.t0 = t0;
  A.t1 = t1;
  A.h = h;
  A.n_traj = n;
  A.y0 = y0;
  A.params = params;
  A.t_final = t_final;
  A.y_final = y_final;
  A.steps = steps;
  A.rhs_evals = rhs_evals;
  A.status = status;
  A.trace_every = (trace_t && trace_capacity > 0) ? trace_every : 0;
  A.trace_capacity = trace_capacity;
  A.trace_t = trace_t;
  A.trace_y = trace_y;
  A.n_samples = n_samples;
  const bool trace = A.trace_every > 0;
  // members per thread and block size from the ensemble size (host/launch_geometry.hpp)
  int n_sm = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const bool two_ok = S->info.n_state * (rk.S + 3) <= 40;
  dfb_host::GeometryCandidate cand[6];
  int nc = 0;
  for (int tpt = two_ok ? 2 : 1; tpt >= 1; --tpt)
    for (int b : {128, 64, 32}) cand[nc++] = {b, tpt, 0, tpt == 2 ? 1.0 : 1.012};
  const dfb_host::LaunchGeometry g = dfb_host::choose_geometry(n, n_sm, cand, nc);
  const std::string ns = ns_of(ARITH);
  // exact sparsity of THIS tableau: structurally-zero entries cost no instruction
  const std::string expr = ns + "::ode_fixed_kernel<" + ns + "::" + S->struct_name + ", " + std::to_string(rk.S) + ", " +
                           hex64(amask) + ", " + std::to_string(bmask) + "u, " + (ARITH == 1 ? "true" : "false") +
                           ", 128, " + std::to_string(g.tpt) + ", " + (trace ? "true" : "false") + ">";
  CUfunction fn = get_kernel(*S, ARITH, "dev/ode_fixed.cuh", expr);
  if (!fn) return int(cudaErrorUnknown);
  void* args[] = {&A};
  const CUresult rc = api(true)->launchKernel(fn, unsigned(g.grid), 1, 1, unsigned(g.block), 1, 1, 0, stream, args, nullptr);
  if (rc != CUDA_SUCCESS) {
    g_error = "cuLaunchKernel failed for " + expr;
    return int(cudaErrorLaunchFailure);
  }
  return 0;
}

template <int ARITH>
int nvrtc_general(const DevProblem& P, int system_id, bool has_loc, const LaunchCfg& c, int locate_batch_threshold) {
  const SourceSys* S = find_source(system_id);
  if (!S) return -1;
  const std::string ns = ns_of(ARITH);
  const std::string expr = ns + "::integrate_general_kernel<" + ns + "::" + S->struct_name + ", " + std::to_string(P.rk.S) +
                           ", " + std::to_string(P.rk.I) + ", " + (has_loc ? "true" : "false") + ", false, " +
                           (P.ode_history ? "true" : "false") + ">";
  CUfunction fn = get_kernel(*S, ARITH, "dev/integrator.cuh", expr);
  if (!fn) return int(cudaErrorUnknown);
  DevProblem Pc = P;
  int batch = locate_batch_threshold, wait = c.locate_max_wait;
  void* args[] = {&Pc, &batch, &wait};
  const CUresult rc = api(true)->launchKernel(fn, unsigned(c.grid), 1, 1, unsigned(c.block), 1, 1, 0, c.stream, args, nullptr);
  if (rc != CUDA_SUCCESS) {
    g_error = "cuLaunchKernel failed for " + expr;
    return int(cudaErrorLaunchFailure);
  }
  return 0;
}

// static facts of the functor, read from the PTX of a one-array probe (no GPU needed to register)
bool probe(SourceSys* S) {
  const std::string main_text =
      "#include \"dev/systems.cuh\"\n"
      "namespace dfb_probe_ns { using Sys = dfb_fast::" + S->struct_name + ";\n"
      "extern \"C\" __device__ const int dfb_probe_info[10] = {Sys::N, Sys::NP, Sys::NA, Sys::ND, Sys::NC,\n"
      "  Sys::uses_history ? 1 : 0, Sys::reads_d ? 1 : 0, Sys::pure_rhs ? 1 : 0, Sys::has_const_delay ? 1 : 0, 12345};\n"
      "extern \"C\" __global__ void dfb_probe_kernel(int* o) { for (int i = 0; i < 10; ++i) o[i] = dfb_probe_info[i]; } }\n";
  std::vector<char> ptx;
  std::string unused;
  if (!compile(*S, 1, main_text, "", 1, &ptx, &unused)) return false;
  const std::string text(ptx.begin(), ptx.end());
  const size_t at = text.find("dfb_probe_info");
  const size_t lb = at == std::string::npos ? at : text.find('{', at);
  const size_t rb = lb == std::string::npos ? lb : text.find('}', lb);
  if (rb == std::string::npos) {
    g_error = "could not read the functor's static facts from the NVRTC probe";
    return false;
  }
  // NVRTC prints the array either as 10 .b32 words or as 40 .b8 bytes (little endian, trailing zeros cut)
  const bool bytes = text.rfind(".b8", at) != std::string::npos && text.rfind(".b8", at) > text.rfind('\n', at);
  long raw[40] = {0};
  int k = 0;
  const char* p = text.c_str() + lb + 1;
  const char* end = text.c_str() + rb;
  while (p < end && k < 40) {
    char* q = nullptr;
    const long x = std::strtol(p, &q, 10);
    if (q == p) {
      ++p;
      continue;
    }
    raw[k++] = x;
    p = q;
  }
  int v[10] = {0};
  for (int i = 0; i < 10; ++i)
    v[i] = bytes ? int(unsigned(raw[4 * i]) | (unsigned(raw[4 * i + 1]) << 8) | (unsigned(raw[4 * i + 2]) << 16) |
                       (unsigned(raw[4 * i + 3]) << 24))
                 : int(raw[i]);
  if (v[9] != 12345) {
    g_error = "unexpected NVRTC probe output";
    return false;
  }
  S->info = dfb_nvrtc_sysinfo{v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]};
  return true;
}

}  // namespace

int dfb_nvrtc_register(const char* source, const char* struct_name, int system_id, dfb_nvrtc_sysinfo* info,
                       DfbLaunchTable* fast, DfbLaunchTable* strict) {
  SourceSys* S = new SourceSys();
  S->system_id = system_id;
  S->source = source;
  S->struct_name = struct_name;
  if (!probe(S)) {
    delete S;
    return -1;
  }
  sources().push_back(S);
  *info = S->info;
  const bool hot = !S->info.uses_history && !S->info.reads_d;
  *fast = DfbLaunchTable{hot ? nvrtc_ode_fixed_shim<1> : nullptr, nullptr, nullptr, nullptr, nullptr, {nvrtc_general<1>}, 1,
                         nullptr, nullptr};
  *strict = DfbLaunchTable{hot ? nvrtc_ode_fixed_shim<0> : nullptr, nullptr, nullptr, nullptr, nullptr, {nvrtc_general<0>}, 1,
                           nullptr, nullptr};
  return 0;
}

const char* dfb_nvrtc_last_error() { return g_error.c_str(); }

void dfb_nvrtc_get_stats(unsigned* compiled, unsigned* cache_hits, double* total_ms, double* last_ms) {
  if (compiled) *compiled = g_stats.compiled;
  if (cache_hits) *cache_hits = g_stats.cache_hits;
  if (total_ms) *total_ms = g_stats.total_ms;
  if (last_ms) *last_ms = g_stats.last_ms;
}

// Build (or fetch from the cache) one instantiation without launching it: the "does it build" check of a
// CPU-only box, and a way to pay the compile time before the first run.
int dfb_nvrtc_prebuild_ode_fixed(int system_id, int arith, const dfb_tableau* rk, int tpt, int trace) {
  const SourceSys* S = find_source(system_id);
  if (!S) {
    g_error = "not a source-registered system";
    return -1;
  }
  unsigned long long amask = 0;
  unsigned bmask = 0;
  for (int i = 0; i < rk->S; ++i) {
    for (int j = 0; j < i; ++j)
      if (rk->a[i * DFB_MAX_STAGES + j] != 0.0) amask |= dfb_fast::amask_bit(i, j);
    if (rk->b[i] != 0.0) bmask |= 1u << i;
  }
  const std::string ns = ns_of(arith);
  const std::string expr = ns + "::ode_fixed_kernel<" + ns + "::" + S->struct_name + ", " + std::to_string(rk->S) + ", " +
                           hex64(amask) + ", " + std::to_string(bmask) + "u, " + (arith == 1 ? "true" : "false") + ", 128, " +
                           std::to_string(tpt) + ", " + (trace ? "true" : "false") + ">";
  std::vector<char> cubin;
  std::string lowered;
  return compile(*S, arith, "#include \"dev/ode_fixed.cuh\"\n", expr, 0, &cubin, &lowered) && !cubin.empty() ? 0 : -1;
}
