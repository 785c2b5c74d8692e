// collective.cu — hsr_allreduce_moments: the one collective of the path (the global fit's moment sum) as a thin
// wrapper on ncclAllReduce, for hosts that bring their own ncclComm_t.  NCCL is resolved with dlopen at the first
// call (the copy already mapped into the process wins: torch ships its own), so libhsr_b200.so carries no
// link-time dependency on it.  The fused peer-memory exchange (poly.cu) is the default inside one node.
#include <dlfcn.h>

#include "hsr_common.cuh"

namespace hsr {

namespace {

// the two NCCL enumerators used here (nccl.h: ncclDataType_t / ncclRedOp_t), stable since NCCL 2.0
constexpr int NCCL_DOUBLE = 8;  // ncclFloat64
constexpr int NCCL_SUM = 0;     // ncclSum

using allreduce_fn = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
using errstr_fn = const char* (*)(int);

struct Nccl {
    allreduce_fn all_reduce = nullptr;
    errstr_fn err_string = nullptr;
    bool tried = false;
};

Nccl& nccl() {
    static Nccl n;
    if (n.tried) return n;
    static const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names)  // the copy the process already loaded (e.g. torch's bundled one)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD)) != nullptr) break;
    if (!h)
        for (const char* nm : names)
            if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (h) {
        n.all_reduce = reinterpret_cast<allreduce_fn>(dlsym(h, "ncclAllReduce"));
        n.err_string = reinterpret_cast<errstr_fn>(dlsym(h, "ncclGetErrorString"));
    }
    __atomic_thread_fence(__ATOMIC_RELEASE);
    n.tried = true;
    return n;
}

}  // namespace

int allreduce_moments_impl(double* moments, long long count, void* comm, cudaStream_t stream) {
    HSR_REQUIRE(moments && comm, HSR_EINVAL, "null moments / communicator");
    HSR_REQUIRE(count >= 0, HSR_EINVAL, "count = %lld", count);
    if (count == 0) return HSR_OK;
    Nccl& n = nccl();
    HSR_REQUIRE(n.all_reduce, HSR_ENCCL, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "no ncclAllReduce");
    const int rc = n.all_reduce(moments, moments, (size_t)count, NCCL_DOUBLE, NCCL_SUM, comm, stream);
    HSR_REQUIRE(rc == 0, HSR_ENCCL, "ncclAllReduce failed: %d (%s)", rc, n.err_string ? n.err_string(rc) : "?");
    return HSR_OK;
}

}  // namespace hsr
