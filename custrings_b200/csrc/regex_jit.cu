// Run-time specialisation of the item-buffered chain kernel (regex_chain_item.cuh) for the plan at hand.
//
// The ahead-of-time build carries a generic instantiation per (steps, classes) that interprets the plan's flags from the
// kernel parameter block, plus a handful of shape specialisations (PlanLit<1..7>).  Here ANY chain gets the same treatment:
// a PlanLit<100> whose every flag, class index and class atom is a literal of the pattern is generated as text, the kernel
// headers (embedded in the library at build time) are compiled with NVRTC for sm_100a (~0.6 s, cached per plan and process)
// and the cubin is loaded through the CUDA runtime's library API.  NVRTC is bound with dlopen: without it (or on any
// failure) the caller falls back to the ahead-of-time kernels.  No tracing, no IR: it is the same hand-written CUDA source.
#include "regex_bits.h"
#include "regex_bits_plan.h"
#include "regex_vm.cuh"
#include <dlfcn.h>
#include <nvrtc.h>
#include <cstddef>
#include <map>
#include <mutex>
#include <sstream>

namespace custr {
namespace bits {

#include "regex_bits_dev.cuh"
#define CUSTR_NO_LAUNCHERS
#define CUSTR_NO_ITEM_LAUNCHERS
#include "regex_chain.cuh"
#include "regex_chain64.cuh"
#include "regex_chain_item.cuh"

struct JitHeader { const char* name; const char* text; };
static const JitHeader k_jit_headers[] = {
#include "../_build/jit_headers.inc"
};

int g_jit_mode = 1;                       // 0 never, 1 large columns without an ahead-of-time shape, 2 always (tests)
static thread_local std::string g_jit_note;  // why the last request was not served (custr_jit_note)
const char* jit_last_note() { return g_jit_note.c_str(); }

namespace {
struct Nvrtc {
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    bool ok = false;
};
const Nvrtc& nvrtc()
{
    static Nvrtc api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"})
            if ((h = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) return;
#define BIND(f) api.f = (decltype(api.f))dlsym(h, "nvrtc" #f)
        BIND(CreateProgram); BIND(CompileProgram); BIND(GetCUBINSize); BIND(GetCUBIN); BIND(GetProgramLogSize); BIND(GetProgramLog); BIND(DestroyProgram);
#undef BIND
        api.ok = api.CreateProgram && api.CompileProgram && api.GetCUBINSize && api.GetCUBIN && api.GetProgramLogSize && api.GetProgramLog && api.DestroyProgram;
    });
    return api;
}

// text of `static constexpr uint32_t name(int a[, int b])` answering from a table
std::string table1(const char* name, const std::vector<uint32_t>& v)
{
    std::ostringstream o;
    o << "    __device__ static constexpr uint32_t " << name << "(int i) { return ";
    for (size_t i = 0; i < v.size(); ++i) o << "i == " << i << " ? " << v[i] << "u : ";
    o << "0u; }\n";
    return o.str();
}
std::string plan_view_text(const ChainDev& cd)
{
    bool opt = false;
    for (uint32_t s = 0; s < cd.nsteps; ++s) opt = opt || cd.steps[s].opt || ((cd.steps[s].exit != 0) != (s + 1 == cd.nsteps));
    std::vector<uint32_t> cls, loop, sopt, sexit, cb, cn, cneg;
    for (uint32_t s = 0; s < cd.nsteps; ++s) {
        cls.push_back(cd.steps[s].cls);
        loop.push_back(cd.steps[s].loop ? 1 : 0);
        sopt.push_back(cd.steps[s].opt ? 1 : 0);
        sexit.push_back(cd.steps[s].exit ? 1 : 0);
    }
    for (uint32_t k = 0; k < cd.nclasses; ++k) {
        cb.push_back(cd.classes[k].builtins);
        cn.push_back(cd.classes[k].natoms);
        cneg.push_back(cd.classes[k].negate ? 1 : 0);
    }
    std::ostringstream o;
    o << "template <> struct PlanLit<100> : PlanLitBase {\n"
      << "    static constexpr bool on = true, opt = " << (opt ? "true" : "false") << ", jit = true;\n"
      << "    static constexpr uint32_t needs = " << cd.needs << "u, end_mask = " << cd.end_mask << "u, before0 = " << cd.steps[0].before
      << "u, builtins = " << cd.builtin_union << "u;\n"
      << "    __device__ static constexpr uint32_t anchored_() { return " << (cd.anchored ? 1 : 0) << "u; }\n"
      << "    __device__ static constexpr uint32_t nclasses_() { return " << cd.nclasses << "u; }\n"
      << table1("step_cls", cls) << table1("step_loop", loop) << table1("step_opt", sopt) << table1("step_exit", sexit)
      << table1("cls_builtins", cb) << table1("cls_natoms", cn) << table1("cls_negate", cneg);
    for (const char* field : {"atom_kind", "atom_lo", "atom_hi"}) {
        o << "    __device__ static constexpr uint32_t " << field << "(int k, int a) { return ";
        for (uint32_t k = 0; k < cd.nclasses; ++k)
            for (uint32_t a = 0; a < cd.classes[k].natoms; ++a) {
                const AtomD& at = cd.classes[k].atoms[a];
                const uint32_t v = field[5] == 'k' ? at.kind : (field[5] == 'l' ? at.lo : at.hi);
                o << "(k == " << k << " && a == " << a << ") ? " << v << "u : ";
            }
        o << "0u; }\n";
    }
    o << "};\n";
    return o.str();
}
std::string escape_macro(const std::string& text)  // a multi-line macro body
{
    std::string r;
    for (char c : text) {
        if (c == '\n') r += " \\\n";
        else r.push_back(c);
    }
    return r;
}

struct JitKernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    bool failed = false;
};
std::mutex g_mu;
std::map<std::string, JitKernel> g_cache;

bool compile(const std::string& view, int ns, int ncls, JitKernel& out, std::string& why)
{
    const Nvrtc& rt = nvrtc();
    std::ostringstream src;
    src << "#define CUSTR_JIT 1\n#define CUSTR_JIT_NS " << ns << "\n#define CUSTR_JIT_NCLS " << ncls << "\n#define CUSTR_NO_LAUNCHERS\n"
        << "#define CUSTR_JIT_PLANLIT " << escape_macro(view) << "\n"
        << "#include \"device_utils.cuh\"\n"
        << "namespace custr { namespace rx { enum : int32_t { CB_W = 1, CB_S = 2, CB_D = 4, CB_NW = 8, CB_NS = 16, CB_ND = 32 }; } }\n"
        << "#include \"regex_bits_plan.h\"\nnamespace custr {\nnamespace bits {\n#include \"regex_bits_dev.cuh\"\n#include \"regex_chain.cuh\"\n"
        << "#include \"regex_chain64.cuh\"\n#include \"regex_chain_item.cuh\"\n}\n}\n";
    std::vector<const char*> names, texts;
    for (const JitHeader& h : k_jit_headers) { names.push_back(h.name); texts.push_back(h.text); }
    nvrtcProgram prog = nullptr;
    const std::string s = src.str();
    if (rt.CreateProgram(&prog, s.c_str(), "custr_jit_chain_item.cu", (int)names.size(), texts.data(), names.data()) != NVRTC_SUCCESS) {
        why = "nvrtcCreateProgram failed";
        return false;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    const nvrtcResult rc = rt.CompileProgram(prog, 4, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t n = 0;
        rt.GetProgramLogSize(prog, &n);
        std::string log(n, ' ');
        if (n) rt.GetProgramLog(prog, &log[0]);
        why = "NVRTC: " + log.substr(0, 600);
        rt.DestroyProgram(&prog);
        return false;
    }
    size_t n = 0;
    rt.GetCUBINSize(prog, &n);
    std::vector<char> cubin(n);
    rt.GetCUBIN(prog, cubin.data());
    rt.DestroyProgram(&prog);
    cudaError_t e = cudaLibraryLoadData(&out.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&out.kernel, out.lib, "custr_jit_chain_item");
    if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)out.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ITEM_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)out.kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
        why = std::string("loading the compiled kernel: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return false;
    }
    return true;
}
}  // namespace

// Launches the plan's own compiled kernel; false = not served (g_jit_note says why), the caller uses the ahead-of-time one.
bool jit_launch_chain_item(const ChainDev& cd, const Args& a, int blocks)
{
    g_jit_note.clear();
    if (!nvrtc().ok) { g_jit_note = "libnvrtc not found"; return false; }
    for (uint32_t k = 0; k < cd.nclasses; ++k)
        if ((cd.classes[k].na_kind == NA_CLASS || cd.classes[k].na_kind == NA_NCLASS) && !cd.classes[k].na_inline) {
            g_jit_note = "a class has no inline definition for non-ASCII characters";
            return false;
        }
    const int ns = (int)cd.nsteps, ncls = cd.nclasses <= 1 ? 1 : (cd.nclasses == 2 ? 2 : (cd.nclasses <= 4 ? 4 : 8));
    const std::string view = plan_view_text(cd);
    int dev = 0;
    cudaGetDevice(&dev);
    const std::string key = std::to_string(dev) + "/" + std::to_string(ns) + "/" + std::to_string(ncls) + "/" + view;
    JitKernel k;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = g_cache.find(key);
        if (it == g_cache.end()) {
            JitKernel fresh;
            std::string why;
            if (!compile(view, ns, ncls, fresh, why)) {
                fresh.failed = true;
                g_jit_note = why;
            }
            it = g_cache.emplace(key, fresh).first;
        }
        k = it->second;
    }
    if (k.failed) {
        if (g_jit_note.empty()) g_jit_note = "compilation failed earlier for this plan";
        return false;
    }
    void* args[] = {(void*)&cd, (void*)&a};
    CUSTR_CUDA(cudaLaunchKernel((const void*)k.kernel, dim3(blocks), dim3(THREADS), args, ITEM_SMEM_BYTES, g_stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return true;
}

}  // namespace bits
}  // namespace custr
