// Runtime specialisation of the single-person kernel for one camera rig (NVRTC + driver API, both loaded with
// dlopen so that the library has no link-time dependency on them).  The generated translation unit defines the
// rig's constants and the batch shape as macros and includes snowtri_p1.cuh from the package's csrc/ directory;
// see the P1_JIT block there.  Every failure is soft: the caller falls back to the precompiled kernel and the
// reason is kept in the handle (snowtri_jit_status).
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "snowtri_internal.h"

namespace {

typedef int (*nvrtcCreateProgram_t)(void**, const char*, const char*, int, const char* const*, const char* const*);
typedef int (*nvrtcCompileProgram_t)(void*, int, const char* const*);
typedef int (*nvrtcGetSize_t)(void*, size_t*);
typedef int (*nvrtcGetData_t)(void*, char*);
typedef int (*nvrtcDestroyProgram_t)(void**);
typedef int (*cuModuleLoadData_t)(void**, const void*);
typedef int (*cuModuleGetFunction_t)(void**, void*, const char*);
typedef int (*cuModuleUnload_t)(void*);
typedef int (*cuFuncSetAttribute_t)(void*, int, int);
typedef int (*cuLaunchKernel_t)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**,
                                void**);

struct Api {
    bool tried = false, ok = false;
    char why[256] = "";
    nvrtcCreateProgram_t create;
    nvrtcCompileProgram_t compile;
    nvrtcGetSize_t cubin_size, log_size;
    nvrtcGetData_t cubin, log;
    nvrtcDestroyProgram_t destroy;
    cuModuleLoadData_t load;
    cuModuleGetFunction_t getfn;
    cuModuleUnload_t unload;
    cuFuncSetAttribute_t setattr;
    cuLaunchKernel_t launch;
} g_api;

void* open_first(const char* const* names) {
    for (int i = 0; names[i]; ++i)
        if (void* h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL)) return h;
    return nullptr;
}

bool load_api() {
    if (g_api.tried) return g_api.ok;
    g_api.tried = true;
    static const char* nv[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                               "/usr/local/cuda/lib64/libnvrtc.so", nullptr};
    static const char* cu[] = {"libcuda.so.1", "libcuda.so", nullptr};
    void* hn = open_first(nv);
    void* hc = open_first(cu);
    if (!hn || !hc) {
        snprintf(g_api.why, sizeof(g_api.why), "%s not found", !hn ? "libnvrtc" : "libcuda");
        return false;
    }
#define SYM(field, lib, name)                                                                   \
    if (!(g_api.field = (decltype(g_api.field))dlsym(lib, name))) {                             \
        snprintf(g_api.why, sizeof(g_api.why), "symbol %s missing", name);                      \
        return false;                                                                           \
    }
    SYM(create, hn, "nvrtcCreateProgram") SYM(compile, hn, "nvrtcCompileProgram") SYM(cubin_size, hn, "nvrtcGetCUBINSize")
    SYM(cubin, hn, "nvrtcGetCUBIN") SYM(log_size, hn, "nvrtcGetProgramLogSize") SYM(log, hn, "nvrtcGetProgramLog")
    SYM(destroy, hn, "nvrtcDestroyProgram") SYM(load, hc, "cuModuleLoadData") SYM(getfn, hc, "cuModuleGetFunction")
    SYM(unload, hc, "cuModuleUnload") SYM(setattr, hc, "cuFuncSetAttribute") SYM(launch, hc, "cuLaunchKernel")
#undef SYM
    g_api.ok = true;
    return true;
}

struct Entry {
    uint64_t key;
    void* mod;
    void* fn;
    int smem_attr;
};
struct Cache {
    Entry e[8];
    int n;
};

uint64_t fnv1a(const std::string& s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
    return h;
}

// FNV-1a of the two headers the generated translation unit includes, as they were when the library was built
// (snowtri_build_stamp.h is written by snowmocap_b200/build.py).  A stale or edited csrc/ next to the library would
// give the JIT kernel another P1Args layout than the host code fills in.
#include "snowtri_build_stamp.h"
bool headers_match_build(const std::string& dir, char* why, size_t n) {
    uint64_t hsh = 1469598103934665603ull;
    const char* files[] = {"/snowtri_math.cuh", "/snowtri_p1.cuh"};
    for (const char* f : files) {
        FILE* fp = fopen((dir + f).c_str(), "rb");
        if (!fp) {
            snprintf(why, n, "failed: %s%s not readable", dir.c_str(), f);
            return false;
        }
        int ch;
        while ((ch = fgetc(fp)) != EOF) hsh = (hsh ^ (unsigned char)ch) * 1099511628211ull;
        fclose(fp);
    }
    if (hsh != SNOWTRI_P1_HEADERS_FNV) {
        snprintf(why, n, "failed: csrc/snowtri_p1.cuh or snowtri_math.cuh differs from the build of this library");
        return false;
    }
    return true;
}

std::string csrc_dir() {
    Dl_info info;
    if (!dladdr((void*)&snowtri_version, &info) || !info.dli_fname) return "";
    std::string p(info.dli_fname);
    const size_t slash = p.rfind('/');
    return (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/csrc";
}

}  // namespace

void snowtri_jit_free(snowtri_t* h) {
    Cache* c = (Cache*)h->jit_cache;
    if (!c) return;
    if (g_api.ok) {
        cudaDeviceSynchronize();  // a kernel of these modules may still be running on a caller stream
        for (int i = 0; i < c->n; ++i) g_api.unload(c->e[i].mod);
    }
    free(c);
    h->jit_cache = nullptr;
}

bool snowtri_jit_cached(const snowtri_t* h, const std::string& source) {
    const Cache* c = (const Cache*)h->jit_cache;
    if (!c) return false;
    const uint64_t key = fnv1a(source);
    for (int i = 0; i < c->n; ++i)
        if (c->e[i].key == key) return true;
    return false;
}

// Compile `source` (or fetch it from the handle's cache) and return the kernel `name`; nullptr on failure.
void* snowtri_jit_get(snowtri_t* h, const std::string& source, const char* name, size_t smem) {
    if (!load_api()) {
        snprintf(h->jit_status, sizeof(h->jit_status), "unavailable: %s", g_api.why);
        return nullptr;
    }
    if (!h->jit_cache) h->jit_cache = calloc(1, sizeof(Cache));
    Cache* c = (Cache*)h->jit_cache;
    const uint64_t key = fnv1a(source);
    Entry* hit = nullptr;
    for (int i = 0; i < c->n; ++i)
        if (c->e[i].key == key) hit = &c->e[i];
    if (!hit) {
        const std::string dir = csrc_dir();
        if (dir.empty()) {
            snprintf(h->jit_status, sizeof(h->jit_status), "failed: cannot locate csrc/ next to the library");
            return nullptr;
        }
        if (!headers_match_build(dir, h->jit_status, sizeof(h->jit_status))) return nullptr;
        void* prog = nullptr;
        if (g_api.create(&prog, source.c_str(), "snowtri_p1_jit.cu", 0, nullptr, nullptr) != 0) {
            snprintf(h->jit_status, sizeof(h->jit_status), "failed: nvrtcCreateProgram");
            return nullptr;
        }
        const std::string inc = "-I" + dir;
        const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo", inc.c_str()};
        const int rc = g_api.compile(prog, 5, opts);
        if (rc != 0) {
            size_t n = 0;
            g_api.log_size(prog, &n);
            std::string log(n + 1, '\0');
            if (n) g_api.log(prog, &log[0]);
            snprintf(h->jit_status, sizeof(h->jit_status), "failed: nvrtc rc=%d: %.400s", rc, log.c_str());
            g_api.destroy(&prog);
            return nullptr;
        }
        size_t nb = 0;
        g_api.cubin_size(prog, &nb);
        std::string cubin(nb, '\0');
        g_api.cubin(prog, &cubin[0]);
        g_api.destroy(&prog);
        void *mod = nullptr, *fn = nullptr;
        int e = g_api.load(&mod, cubin.data());
        if (e == 0) e = g_api.getfn(&fn, mod, name);
        if (e != 0) {
            snprintf(h->jit_status, sizeof(h->jit_status), "failed: module load / lookup, CUresult %d", e);
            if (mod) g_api.unload(mod);
            return nullptr;
        }
        if (c->n == 8) {  // cache full: drop the oldest -- after every launch of it has finished
            cudaDeviceSynchronize();
            g_api.unload(c->e[0].mod);
            memmove(&c->e[0], &c->e[1], 7 * sizeof(Entry));
            c->n = 7;
        }
        hit = &c->e[c->n++];
        hit->key = key; hit->mod = mod; hit->fn = fn; hit->smem_attr = 0;
        snprintf(h->jit_status, sizeof(h->jit_status), "compiled (%zu byte cubin)", nb);
    }
    if ((int)smem > 48 * 1024 && hit->smem_attr < (int)smem) {
        if (g_api.setattr(hit->fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)smem) != 0) {
            snprintf(h->jit_status, sizeof(h->jit_status), "failed: cuFuncSetAttribute(%zu B)", smem);
            return nullptr;
        }
        hit->smem_attr = (int)smem;
    }
    return hit->fn;
}

int snowtri_jit_launch(void* fn, int grid, int block, size_t smem, void* stream, void* args) {
    void* params[] = {args};
    return g_api.launch(fn, (unsigned)grid, 1, 1, (unsigned)block, 1, 1, (unsigned)smem, stream, params, nullptr);
}
