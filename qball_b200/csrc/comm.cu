// qball_b200/csrc/comm.cu -- the path's two collectives behind the C ABI:
//   qb200_allreduce_rho      <->  wfcontext->dsum('r', np012loc, 1, &rhor[ispin][0], np012loc)      ChargeDensity.cc:309
//   qb200_allreduce_scalars  <->  ctxt_.dsum('r',1,1,&enl,1) NonLocalPotential.cc:2629; psi.wfcontext()->dsum(14,...)
//                                 EnergyFunctional.cc:1294; the dsum of nelectrons ChargeDensity.cc:528
// over the GPUs of one box (band parallelism: nprow = 1, one rank per GPU), with NCCL over NVLink / NVSwitch INSIDE the
// library: a C++ caller (the reference's shim) needs nothing but an out-of-band broadcast of the 128-byte unique id (MPI_Bcast).
// NCCL is bound at run time (dlopen of libnccl.so.2): single-GPU users carry no NCCL dependency, and inside a process that
// already holds an NCCL (torch's) that copy is the one used.
#include "qb200_internal.h"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <mutex>

struct qb200_comm {
  int device, rank, nranks;
  ncclComm_t comm;
  cudaStream_t stream;         // the library's own stream for host-pointer calls
  double* stage; size_t stage_cap;
};

namespace {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
std::mutex g_mu;

int nccl_load()
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_nccl.h) return QB200_OK;
  const char* names[] = { "libnccl.so.2", "libnccl.so" };
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) { qb200::set_error(std::string("qb200_comm: cannot load NCCL: ") + dlerror()); return QB200_EUNSUPPORTED; }
  NcclApi a;
  a.h = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString) {
    qb200::set_error("qb200_comm: NCCL library lacks a required symbol"); return QB200_EUNSUPPORTED;
  }
  g_nccl = a;
  return QB200_OK;
}

int nccl_fail(ncclResult_t r, const char* what)
{
  qb200::set_error(std::string("qb200_comm: ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
  return QB200_ECUDA;
}
#define QB_NCCL(x) do { ncclResult_t r__ = (x); if (r__ != ncclSuccess) return nccl_fail(r__, #x); } while (0)

}  // namespace

using namespace qb200;

static_assert(sizeof(ncclUniqueId) == QB200_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");

extern "C" int qb200_comm_get_unique_id(void* id)
{
  if (!id) { set_error("qb200_comm_get_unique_id: bad argument"); return QB200_EINVAL; }
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId u;
  QB_NCCL(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return QB200_OK;
}

extern "C" int qb200_comm_init(qb200_comm** out, int device, const void* id, int rank, int nranks)
{
  if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) { set_error("qb200_comm_init: bad argument"); return QB200_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  QB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_error("qb200_comm_init: no such CUDA device"); return QB200_ENODEV; }
  int rc = nccl_load();
  if (rc) return rc;
  QB_CUDA(cudaSetDevice(device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclComm_t c;
  QB_NCCL(g_nccl.CommInitRank(&c, nranks, u, rank));
  qb200_comm* q = new qb200_comm();
  q->device = device; q->rank = rank; q->nranks = nranks; q->comm = c; q->stream = nullptr; q->stage = nullptr; q->stage_cap = 0;
  const cudaError_t e = cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { g_nccl.CommDestroy(c); delete q; return cuda_fail(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__); }
  *out = q;
  return QB200_OK;
}

extern "C" int qb200_comm_destroy(qb200_comm* q)
{
  if (!q) return QB200_OK;
  cudaSetDevice(q->device);
  if (q->stream) { cudaStreamSynchronize(q->stream); cudaStreamDestroy(q->stream); }
  if (q->stage) cudaFree(q->stage);
  if (g_nccl.CommDestroy) g_nccl.CommDestroy(q->comm);
  delete q;
  return QB200_OK;
}

extern "C" long long qb200_comm_query(const qb200_comm* q, int what)
{
  if (!q) return -1;
  switch (what) {
    case 0: return q->rank;
    case 1: return q->nranks;
    case 2: { int v = 0; if (g_nccl.GetVersion) g_nccl.GetVersion(&v); return v; }
    default: return -1;
  }
}

// sum of n doubles over the ranks, in place.  Device pointer: enqueued on `stream` (cudaStream_t as void*; NULL = the legacy
// default stream), asynchronous like a kernel launch.  Host pointer: staged through the communicator's buffer and stream,
// synchronous.  Every rank must pass the same n.
static int allreduce_impl(qb200_comm* q, double* x, long long n, void* stream)
{
  if (!q || !x || n < 0) { set_error("qb200_allreduce: bad argument"); return QB200_EINVAL; }
  if (n == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(q->device));
  if (is_device_ptr(x)) {
    if (q->nranks > 1) QB_NCCL(g_nccl.AllReduce(x, x, (size_t)n, ncclDouble, ncclSum, q->comm, (cudaStream_t)stream));
    return QB200_OK;
  }
  if (q->nranks == 1) return QB200_OK;
  if (q->stage_cap < (size_t)n) {
    if (q->stage) { cudaFree(q->stage); q->stage = nullptr; q->stage_cap = 0; }
    QB_CUDA(cudaMalloc((void**)&q->stage, (size_t)n * sizeof(double)));
    q->stage_cap = (size_t)n;
  }
  QB_CUDA(cudaMemcpyAsync(q->stage, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, q->stream));
  QB_NCCL(g_nccl.AllReduce(q->stage, q->stage, (size_t)n, ncclDouble, ncclSum, q->comm, q->stream));
  QB_CUDA(cudaMemcpyAsync(x, q->stage, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, q->stream));
  QB_CUDA(cudaStreamSynchronize(q->stream));
  return QB200_OK;
}

extern "C" int qb200_allreduce_rho(qb200_comm* q, double* rho, long long n, void* stream) { return allreduce_impl(q, rho, n, stream); }

extern "C" int qb200_allreduce_scalars(qb200_comm* q, double* vals, int n)
{
  if (q && vals && n > 0 && is_device_ptr(vals)) { set_error("qb200_allreduce_scalars: host array expected"); return QB200_EINVAL; }
  return allreduce_impl(q, vals, n, nullptr);
}
