// Whole-forward native entry ("a bflow_forward_graph that runs the whole captured path").
//
// The launch plan of one (batch, height, width, iterations) is produced once by the Python planner (bflow_b200/export.py records the
// same launch list RAFTSpline.forward replays) and written to a PLAN FILE; this file is everything a native caller needs:
//
//     bflow_forward* f;   bflow_forward_load("d_480x640.plan", &f);
//     bflow_forward_run(f, voxel, NULL, NULL, NULL, low, up, stream);          // host or device pointers
//     bflow_forward_destroy(f);
//
// No Python, no torch at run time.  How the plan stays valid across processes without relocating a single pointer: every device
// allocation of the exporting process (packed weights, workspace, I/O buffers, tensor-map targets) comes from ONE arena that is
// mapped with the CUDA virtual-memory API at a FIXED virtual address (bflow_arena_open: cuMemAddressReserve with the address as a
// requirement, cuMemCreate + cuMemMap behind it; PyTorch allocates from it through a pluggable allocator).  The loader reserves the
// same range, maps fresh memory, restores the non-zero 64 KB chunks (the weight images) and replays the recorded launches -- every
// absolute pointer in every descriptor and inside every encoded tensor map is valid as written.
//
// Plan file (little endian): "BFLOWPLN" u32 version | u64 arena_base, arena_reserve, arena_used | i32 meta[16] |
//   u64 io_off[6] (voxel, img0, img1, init, low, up; ~0 = absent) u64 io_bytes[6] | u32 n_chunks { u64 off, u32 len, bytes } |
//   u32 n_items { u8 kind: 0 launch / 1 fork / 2 join; launch: u8 stream, u16 name_len, name, u16 nargs { u8 k: 0 int64, 1 double,
//   2 pointer value, 3 blob: u32 len + bytes } }
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace bflow {

// ---------------------------------------------------------------------------------------------------------------------
// fixed-address arena on the CUDA virtual-memory API
// ---------------------------------------------------------------------------------------------------------------------
struct Vmm {
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*addr_free)(CUdeviceptr, size_t);
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
    CUresult (*release)(CUmemGenericAllocationHandle);
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*unmap)(CUdeviceptr, size_t);
    CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
    CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
    bool ok;
};
static bool entry(const char* name, void** fn) {
    cudaDriverEntryPointQueryResult q;
    return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn != nullptr;
}
static Vmm& vmm() {
    static Vmm v = {};
    static bool tried = false;
    if (!tried) {
        tried = true;
        cudaFree(nullptr);      // make sure the primary context exists
        v.ok = entry("cuMemAddressReserve", (void**)&v.reserve) && entry("cuMemAddressFree", (void**)&v.addr_free) && entry("cuMemCreate", (void**)&v.create) &&
               entry("cuMemRelease", (void**)&v.release) && entry("cuMemMap", (void**)&v.map) && entry("cuMemUnmap", (void**)&v.unmap) &&
               entry("cuMemSetAccess", (void**)&v.set_access) && entry("cuMemGetAllocationGranularity", (void**)&v.granularity);
    }
    return v;
}

struct Arena {
    CUdeviceptr base = 0;
    size_t reserved = 0, used = 0, gran = 0;
    int device = 0;
    std::vector<CUmemGenericAllocationHandle> handles;
    std::vector<std::pair<size_t, size_t>> maps;      // (offset, size) of every mapping
};
static Arena* g_arena = nullptr;      // the exporting process has exactly one

static int arena_open(Arena& a, unsigned long long base, unsigned long long reserve) {
    Vmm& v = vmm();
    if (!v.ok) {
        set_error("arena: the CUDA virtual-memory API is not available from this driver");
        return BFLOW_ERR_CUDA;
    }
    cudaGetDevice(&a.device);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = a.device;
    if (v.granularity(&a.gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || a.gran == 0) a.gran = 2u << 20;
    a.reserved = (reserve + a.gran - 1) / a.gran * a.gran;
    CUdeviceptr got = 0;
    if (v.reserve(&got, a.reserved, 0, (CUdeviceptr)base, 0) != CUDA_SUCCESS) {
        set_error("arena: cuMemAddressReserve failed");
        return BFLOW_ERR_CUDA;
    }
    if (got != (CUdeviceptr)base) {      // the address is a requirement here, not a hint: every pointer of the plan is absolute
        v.addr_free(got, a.reserved);
        set_error("arena: the plan's virtual address range is not free in this process");
        return BFLOW_ERR_CUDA;
    }
    a.base = got;
    a.used = 0;
    return BFLOW_OK;
}
static void* arena_map(Arena& a, size_t size) {
    Vmm& v = vmm();
    size = (size + a.gran - 1) / a.gran * a.gran;
    if (a.base == 0 || a.used + size > a.reserved) return nullptr;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = a.device;
    CUmemGenericAllocationHandle h;
    if (v.create(&h, size, &prop, 0) != CUDA_SUCCESS) return nullptr;
    if (v.map(a.base + a.used, size, 0, h, 0) != CUDA_SUCCESS) {
        v.release(h);
        return nullptr;
    }
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (v.set_access(a.base + a.used, size, &acc, 1) != CUDA_SUCCESS) {
        v.unmap(a.base + a.used, size);
        v.release(h);
        return nullptr;
    }
    void* p = reinterpret_cast<void*>(a.base + a.used);
    cudaMemset(p, 0, size);      // fresh memory is not zero; the exporter saves only non-zero chunks and torch.empty regions must not look like data
    a.handles.push_back(h);
    a.maps.push_back({a.used, size});
    a.used += size;
    return p;
}
static void arena_close(Arena& a) {
    Vmm& v = vmm();
    if (a.base == 0) return;
    cudaDeviceSynchronize();
    for (size_t i = 0; i < a.maps.size(); ++i) {
        v.unmap(a.base + a.maps[i].first, a.maps[i].second);
        v.release(a.handles[i]);
    }
    v.addr_free(a.base, a.reserved);
    a = Arena();
}

// ---------------------------------------------------------------------------------------------------------------------
// launch replay
// ---------------------------------------------------------------------------------------------------------------------
struct Arg {
    int kind;      // 0 int64, 1 double, 2 pointer
    long long i;
    double f;
    const void* p;
};
struct Item {
    int kind, stream;      // 0 launch, 1 fork, 2 join
    std::string name;
    std::vector<Arg> args;
};

#define AI(k) ((int)a[k].i)
#define AL(k) (a[k].i)
#define AF(k) ((float)a[k].f)
#define AP(k) (const_cast<void*>(a[k].p))
#define FP(k) (reinterpret_cast<const float*>(a[k].p))
#define MF(k) (reinterpret_cast<float*>(const_cast<void*>(a[k].p)))
#define CD(k) (reinterpret_cast<const bflow_conv_desc*>(a[k].p))

static int call(const std::string& n, const std::vector<Arg>& a, void* s) {
    const size_t na = a.size();
#define IS(fn, cnt) (n == #fn && na == (cnt))
    if (IS(bflow_zero, 2)) return bflow_zero(AP(0), (unsigned long long)AL(1), s);
    if (IS(bflow_nchw_to_nhwc, 11)) return bflow_nchw_to_nhwc(FP(0), MF(1), AI(2), AI(3), AI(4), AI(5), AI(6), AI(7), AI(8), AF(9), AF(10), s);
    if (IS(bflow_nhwc_to_nchw, 7)) return bflow_nhwc_to_nchw(FP(0), MF(1), AI(2), AI(3), AI(4), AI(5), AI(6), s);
    if (IS(bflow_conv2d_nhwc, 1)) return bflow_conv2d_nhwc(CD(0), s);
    if (IS(bflow_conv2d_small_n, 1)) return bflow_conv2d_small_n(CD(0), s);
    if (IS(bflow_conv2d_thin7, 1)) return bflow_conv2d_thin7(CD(0), s);
    if (IS(bflow_conv2d_nhwc_tc3, 6)) return bflow_conv2d_nhwc_tc3(CD(0), AP(1), AP(2), AI(3), AF(4), reinterpret_cast<int*>(AP(5)), s);
    if (IS(bflow_conv2d_nhwc_tc3o, 7)) return bflow_conv2d_nhwc_tc3o(CD(0), AP(1), AP(2), AP(3), AI(4), AF(5), reinterpret_cast<int*>(AP(6)), s);
    if (IS(bflow_conv2d_nhwc_tc3s, 7)) return bflow_conv2d_nhwc_tc3s(CD(0), AP(1), AP(2), AI(3), AF(4), AI(5), reinterpret_cast<int*>(AP(6)), s);
    if (IS(bflow_conv2d_slab64, 5)) return bflow_conv2d_slab64(CD(0), AP(1), AP(2), AF(3), reinterpret_cast<int*>(AP(4)), s);
    if (IS(bflow_conv2d_stem7, 9))
        return bflow_conv2d_stem7(CD(0), AP(1), AI(2), reinterpret_cast<const int*>(a[3].p), AI(4), AF(5), AF(6), AF(7), reinterpret_cast<int*>(AP(8)), s);
    if (IS(bflow_im2col_split16, 17))
        return bflow_im2col_split16(FP(0), AI(1), AI(2), AI(3), AI(4), AI(5), AI(6), AI(7), AI(8), AI(9), AI(10), AI(11), AF(12), AF(13), AP(14), AP(15), AI(16), s);
    if (IS(bflow_split_f16, 7)) return bflow_split_f16(FP(0), AI(1), AP(2), AP(3), AI(4), AL(5), AI(6), s);
    if (IS(bflow_pack_b_tc, 8)) return bflow_pack_b_tc(FP(0), AI(1), AP(2), AI(3), AI(4), AI(5), AI(6), AI(7), s);
    if (IS(bflow_plane_sums, 6)) return bflow_plane_sums(FP(0), AI(1), reinterpret_cast<double*>(AP(2)), AI(3), AI(4), AI(5), s);
    if (IS(bflow_instnorm_relu, 12))
        return bflow_instnorm_relu(FP(0), AI(1), reinterpret_cast<const double*>(a[2].p), FP(3), AI(4), reinterpret_cast<const double*>(a[5].p), MF(6), AI(7), AI(8),
                                   AI(9), AI(10), AF(11), s);
    if (IS(bflow_instnorm_relu16, 18))
        return bflow_instnorm_relu16(FP(0), AI(1), reinterpret_cast<const double*>(a[2].p), FP(3), AI(4), reinterpret_cast<const double*>(a[5].p), a[6].p, a[7].p,
                                     AI(8), MF(9), AI(10), AP(11), AP(12), AI(13), AI(14), AI(15), AI(16), AF(17), s);
    if (IS(bflow_corr_volume, 7)) return bflow_corr_volume(FP(0), AI(1), FP(2), MF(3), AI(4), AI(5), AI(6), s);
    if (IS(bflow_corr_pool, 5)) return bflow_corr_pool(FP(0), MF(1), AL(2), AI(3), AI(4), s);
    if (IS(bflow_corr_pool_tiled, 5)) return bflow_corr_pool_tiled(FP(0), MF(1), AL(2), AI(3), AI(4), s);
    if (IS(bflow_corr_lookup, 1)) return bflow_corr_lookup(reinterpret_cast<const bflow_lookup_desc*>(a[0].p), s);
    if (IS(bflow_corr_lookup_otf, 1)) return bflow_corr_lookup_otf(reinterpret_cast<const bflow_lookup_otf_desc*>(a[0].p), s);
    if (IS(bflow_feat_pool, 8)) return bflow_feat_pool(FP(0), MF(1), AI(2), AI(3), AI(4), AI(5), AI(6), AI(7), s);
    if (IS(bflow_cvx_upsample, 11)) return bflow_cvx_upsample(FP(0), AI(1), AI(2), FP(3), AI(4), AI(5), MF(6), AI(7), AI(8), AI(9), AI(10), s);
#undef IS
    set_error("forward: the plan names an entry point (or an argument count) this library does not replay");
    return BFLOW_ERR_INVALID;
}

struct Reader {
    FILE* f;
    bool ok = true;
    template <typename T>
    T get() {
        T v = T();
        if (ok && fread(&v, sizeof(T), 1, f) != 1) ok = false;
        return v;
    }
    void bytes(void* dst, size_t n) {
        if (ok && n && fread(dst, 1, n, f) != n) ok = false;
    }
};

}  // namespace bflow

struct bflow_forward {
    bflow::Arena arena;
    int meta[16];
    unsigned long long io_off[6], io_bytes[6];
    std::vector<bflow::Item> items;
    std::vector<std::vector<unsigned char>*> blobs;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int* err_word = nullptr;
};

static int replay(bflow_forward* f, cudaStream_t main) {
    for (const bflow::Item& it : f->items) {
        if (it.kind == 0) {
            const int rc = bflow::call(it.name, it.args, it.stream == 1 ? f->side : main);
            if (rc != BFLOW_OK) return rc;
        } else if (it.kind == 1) {
            if (cudaEventRecord(f->ev_fork, main) != cudaSuccess || cudaStreamWaitEvent(f->side, f->ev_fork, 0) != cudaSuccess) return BFLOW_ERR_CUDA;
        } else {
            if (cudaEventRecord(f->ev_join, f->side) != cudaSuccess || cudaStreamWaitEvent(main, f->ev_join, 0) != cudaSuccess) return BFLOW_ERR_CUDA;
        }
    }
    return BFLOW_OK;
}

extern "C" void bflow_forward_destroy(bflow_forward* f) {
    if (f == nullptr) return;
    cudaDeviceSynchronize();
    if (f->exec) cudaGraphExecDestroy(f->exec);
    if (f->graph) cudaGraphDestroy(f->graph);
    if (f->side) cudaStreamDestroy(f->side);
    if (f->ev_fork) cudaEventDestroy(f->ev_fork);
    if (f->ev_join) cudaEventDestroy(f->ev_join);
    bflow::arena_close(f->arena);
    for (auto* b : f->blobs) delete b;
    delete f;
}

extern "C" int bflow_forward_load(const char* path, bflow_forward** out) {
    BFLOW_REQUIRE(path != nullptr && out != nullptr, "forward_load: null argument");
    *out = nullptr;
    FILE* fp = fopen(path, "rb");
    BFLOW_REQUIRE(fp != nullptr, "forward_load: cannot open the plan file");
    bflow::Reader r{fp};
    char magic[8];
    r.bytes(magic, 8);
    const unsigned version = r.get<unsigned>();
    if (!r.ok || memcmp(magic, "BFLOWPLN", 8) != 0 || version != 1u) {
        fclose(fp);
        bflow::set_error("forward_load: not a bflow_b200 plan file (version 1)");
        return BFLOW_ERR_INVALID;
    }
    bflow_forward* f = new bflow_forward();
    auto fail = [&](const char* msg, int code) {
        if (msg != nullptr) bflow::set_error(msg);
        fclose(fp);
        bflow_forward_destroy(f);
        return code;
    };
    const unsigned long long base = r.get<unsigned long long>(), reserve = r.get<unsigned long long>(), used = r.get<unsigned long long>();
    r.bytes(f->meta, sizeof(f->meta));
    r.bytes(f->io_off, sizeof(f->io_off));
    r.bytes(f->io_bytes, sizeof(f->io_bytes));
    if (!r.ok || used == 0 || used > reserve) return fail("forward_load: truncated or inconsistent header", BFLOW_ERR_INVALID);
    if (f->meta[15] != BFLOW_ABI_VERSION) return fail("forward_load: the plan was exported by a library with another ABI version", BFLOW_ERR_INVALID);
    int rc = bflow::arena_open(f->arena, base, reserve);
    if (rc != BFLOW_OK) return fail(nullptr, rc);
    for (unsigned long long done = 0; done < used;) {      // fresh device memory behind the plan's addresses, zero-filled
        const unsigned long long piece = used - done < (1ull << 30) ? used - done : (1ull << 30);
        if (bflow::arena_map(f->arena, (size_t)piece) == nullptr) return fail("forward_load: out of device memory while mapping the arena", BFLOW_ERR_CUDA);
        done = f->arena.used;
    }
    const unsigned n_chunks = r.get<unsigned>();
    std::vector<unsigned char> buf;
    for (unsigned c = 0; c < n_chunks && r.ok; ++c) {      // the non-zero chunks: packed weights, biases
        const unsigned long long off = r.get<unsigned long long>();
        const unsigned len = r.get<unsigned>();
        if (!r.ok || off + len > used) return fail("forward_load: bad data chunk", BFLOW_ERR_INVALID);
        buf.resize(len);
        r.bytes(buf.data(), len);
        if (r.ok && cudaMemcpy(reinterpret_cast<void*>(base + off), buf.data(), len, cudaMemcpyHostToDevice) != cudaSuccess)
            return fail("forward_load: copying the weights failed", BFLOW_ERR_CUDA);
    }
    const unsigned n_items = r.get<unsigned>();
    for (unsigned i = 0; i < n_items && r.ok; ++i) {
        bflow::Item it;
        it.kind = r.get<unsigned char>();
        it.stream = 0;
        if (it.kind == 0) {
            it.stream = r.get<unsigned char>();
            const unsigned short nl = r.get<unsigned short>();
            it.name.resize(nl);
            r.bytes(&it.name[0], nl);
            const unsigned short na = r.get<unsigned short>();
            for (unsigned short k = 0; k < na && r.ok; ++k) {
                bflow::Arg a = {0, 0, 0.0, nullptr};
                const unsigned char kind = r.get<unsigned char>();
                if (kind == 0) {
                    a.kind = 0;
                    a.i = r.get<long long>();
                } else if (kind == 1) {
                    a.kind = 1;
                    a.f = r.get<double>();
                } else if (kind == 2) {
                    a.kind = 2;
                    a.p = reinterpret_cast<const void*>(r.get<unsigned long long>());
                } else if (kind == 3) {
                    const unsigned len = r.get<unsigned>();
                    auto* blob = new std::vector<unsigned char>(len + 64);
                    f->blobs.push_back(blob);
                    unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(blob->data()) + 63) & ~(uintptr_t)63);      // tensor maps want 64-byte alignment
                    r.bytes(p, len);
                    a.kind = 2;
                    a.p = p;
                } else {
                    return fail("forward_load: unknown argument kind", BFLOW_ERR_INVALID);
                }
                it.args.push_back(a);
            }
        }
        f->items.push_back(it);
    }
    if (!r.ok) return fail("forward_load: truncated plan file", BFLOW_ERR_INVALID);
    fclose(fp);
    fp = nullptr;
    auto fail2 = [&](const char* msg, int code) {
        if (msg != nullptr) bflow::set_error(msg);
        bflow_forward_destroy(f);
        return code;
    };
    // eager warm-up on a private stream (sets the per-kernel attributes, surfaces contract errors outside the capture), then the CUDA graph
    cudaStream_t main;
    if (cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&f->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&f->ev_join, cudaEventDisableTiming) != cudaSuccess)
        return fail2("forward_load: stream / event creation failed", BFLOW_ERR_CUDA);
    rc = replay(f, main);
    if (rc == BFLOW_OK && cudaStreamSynchronize(main) != cudaSuccess) {
        bflow::set_error("forward_load: the warm-up forward failed on the device");
        rc = BFLOW_ERR_CUDA;
    }
    if (rc == BFLOW_OK) {
        if (cudaStreamBeginCapture(main, cudaStreamCaptureModeThreadLocal) != cudaSuccess) rc = BFLOW_ERR_CUDA;
        if (rc == BFLOW_OK) rc = replay(f, main);
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(main, &g);
        if (rc == BFLOW_OK && (ce != cudaSuccess || g == nullptr)) {
            bflow::set_error("forward_load: graph capture failed");
            rc = BFLOW_ERR_CUDA;
        }
        f->graph = g;
        if (rc == BFLOW_OK && cudaGraphInstantiate(&f->exec, f->graph, 0) != cudaSuccess) {
            bflow::set_error("forward_load: graph instantiation failed");
            rc = BFLOW_ERR_CUDA;
        }
    }
    cudaStreamDestroy(main);
    if (rc != BFLOW_OK) return fail2(nullptr, rc);
    *out = f;
    return BFLOW_OK;
}

extern "C" int bflow_forward_info(const bflow_forward* f, int* info16) {
    BFLOW_REQUIRE(f != nullptr && info16 != nullptr, "forward_info: null argument");
    memcpy(info16, f->meta, sizeof(f->meta));
    return BFLOW_OK;
}

extern "C" int bflow_forward_run(bflow_forward* f, const float* voxel, const float* image0, const float* image1, const float* flow_init, float* low_out,
                                 float* up_out, void* stream) {
    BFLOW_REQUIRE(f != nullptr && f->exec != nullptr, "forward_run: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const float* in[4] = {voxel, image0, image1, flow_init};
    for (int i = 0; i < 4; ++i) {
        const bool has = f->io_off[i] != ~0ull;
        BFLOW_REQUIRE(i == 3 || has == (in[i] != nullptr), "forward_run: pass exactly the inputs the plan was exported with (voxel grid and / or the two images)");
        if (!has) continue;
        void* dst = reinterpret_cast<void*>(f->arena.base + f->io_off[i]);
        cudaError_t e = in[i] != nullptr ? cudaMemcpyAsync(dst, in[i], f->io_bytes[i], cudaMemcpyDefault, st) : cudaMemsetAsync(dst, 0, f->io_bytes[i], st);
        if (e != cudaSuccess) {
            bflow::set_error(cudaGetErrorString(e));
            return BFLOW_ERR_CUDA;
        }
    }
    if (cudaGraphLaunch(f->exec, st) != cudaSuccess) {
        bflow::set_error("forward_run: cudaGraphLaunch failed");
        return BFLOW_ERR_CUDA;
    }
    float* outp[2] = {low_out, up_out};
    for (int i = 0; i < 2; ++i) {
        if (outp[i] == nullptr) continue;
        if (cudaMemcpyAsync(outp[i], reinterpret_cast<const void*>(f->arena.base + f->io_off[4 + i]), f->io_bytes[4 + i], cudaMemcpyDefault, st) != cudaSuccess) {
            bflow::set_error("forward_run: copying the results failed");
            return BFLOW_ERR_CUDA;
        }
    }
    return BFLOW_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// exporter side: the process-wide arena PyTorch allocates from (torch.cuda.memory.CUDAPluggableAllocator)
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int bflow_arena_open(unsigned long long base, unsigned long long reserve_bytes) {
    BFLOW_REQUIRE(bflow::g_arena == nullptr, "arena_open: an arena is already open in this process");
    BFLOW_REQUIRE(base != 0 && reserve_bytes != 0, "arena_open: bad range");
    bflow::Arena* a = new bflow::Arena();
    const int rc = bflow::arena_open(*a, base, reserve_bytes);
    if (rc != BFLOW_OK) {
        delete a;
        return rc;
    }
    bflow::g_arena = a;
    return BFLOW_OK;
}
extern "C" void* bflow_arena_alloc(long size, int device, void* stream) {
    (void)device;
    (void)stream;
    return bflow::g_arena != nullptr && size > 0 ? bflow::arena_map(*bflow::g_arena, (size_t)size) : nullptr;
}
extern "C" void bflow_arena_free(void* ptr, long size, int device, void* stream) {      // the arena is bump-allocated and lives until bflow_arena_close
    (void)ptr;
    (void)size;
    (void)device;
    (void)stream;
}
extern "C" int bflow_arena_read(unsigned long long offset, void* dst_host, unsigned long long bytes) {      // synchronous device -> host copy of an arena range
    BFLOW_REQUIRE(bflow::g_arena != nullptr && dst_host != nullptr && offset + bytes <= bflow::g_arena->used, "arena_read: bad range");
    if (cudaMemcpy(dst_host, reinterpret_cast<const void*>(bflow::g_arena->base + offset), bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
        bflow::set_error("arena_read: cudaMemcpy failed");
        return BFLOW_ERR_CUDA;
    }
    return BFLOW_OK;
}
extern "C" unsigned long long bflow_arena_used(void) { return bflow::g_arena != nullptr ? (unsigned long long)bflow::g_arena->used : 0ull; }
extern "C" int bflow_arena_close(void) {
    if (bflow::g_arena != nullptr) {
        bflow::arena_close(*bflow::g_arena);
        delete bflow::g_arena;
        bflow::g_arena = nullptr;
    }
    return BFLOW_OK;
}
