// libchordb200: one-shot all-reduce of a 12-bin chroma sum over peer GPU memory (NVLink / NVSwitch),
// fused into the last CTA of the producing kernel (he.cu: he2048w_kernel; common.cuh:
// comm_allreduce12).  The path has exactly one exchange step -- the sum of the per-GPU 12-bin
// vectors -- and at 96 bytes it is pure latency: a separate NCCL kernel per step
// costs 13-28 us against a 165 us compute kernel (round 1: 0.86 weak-scaling efficiency at 8
// GPUs).  Here every rank's finalising warp stores its 12 doubles + a flag straight into every
// peer's mailbox (P2P stores), waits for the peers' flags and sums in rank order.
//
// Set-up (host, once): every rank allocates its mailbox with cudaMalloc, exports a CUDA IPC handle
// (cdb_comm_alloc), the ranks exchange the 64-byte handles by any means (the Python side uses
// torch.distributed.all_gather_object) and map each other's mailboxes (cdb_comm_connect).
#include <cstring>

#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == CDB_IPC_HANDLE_BYTES, "IPC handle size");

static size_t mail_bytes(int world) {
  return (size_t)2 * world * CDB_MAIL_STRIDE * sizeof(double) + 64;  // + status word
}

extern "C" {

int cdb_comm_alloc(cdb_handle* h, int world, unsigned char* ipc_handle_out) {
  if (!h || !ipc_handle_out) return CDB_E_NULL;
  if (world < 1 || world > CDB_MAX_PEERS) return cdb_fail(h, CDB_E_INVALID, "world %d", world);
  if (h->comm) return cdb_fail(h, CDB_E_INVALID, "communicator already exists on this handle");
  CDB_CUDA(h, cudaSetDevice(h->device));
  Comm* c = new Comm();
  c->world = world;
  const size_t bytes = mail_bytes(world);
  cudaError_t e = cudaMalloc(&c->local, bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t hd;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, c->local);
  if (e != cudaSuccess) {
    if (c->local) cudaFree(c->local);
    delete c;
    return cdb_fail(h, (int)e, "cdb_comm_alloc: %s", cudaGetErrorString(e));
  }
  std::memcpy(ipc_handle_out, &hd, sizeof(hd));
  c->d_status = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(c->local) + bytes - 64);
  h->comm = c;
  return 0;
}

int cdb_comm_connect(cdb_handle* h, int rank, int world, const unsigned char* all_handles) {
  if (!h || !all_handles) return CDB_E_NULL;
  Comm* c = h->comm;
  if (!c || c->world != world || rank < 0 || rank >= world)
    return cdb_fail(h, CDB_E_INVALID, "cdb_comm_connect: call cdb_comm_alloc(world=%d) first", world);
  CDB_CUDA(h, cudaSetDevice(h->device));
  c->rank = rank;
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      c->mail[q] = reinterpret_cast<double*>(c->local);
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, all_handles + (size_t)q * CDB_IPC_HANDLE_BYTES, sizeof(hd));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return cdb_fail(h, (int)e, "cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e));
    c->mail[q] = reinterpret_cast<double*>(p);
  }
  return 0;
}

// 0 = healthy, 1 = a collective timed out waiting for a peer (its result was NaN), <0 = no communicator
int cdb_comm_status(cdb_handle* h) {
  if (!h) return CDB_E_NULL;
  if (!h->comm) return CDB_E_INVALID;
  int v = 0;
  if (cudaMemcpy(&v, h->comm->d_status, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

int cdb_comm_destroy(cdb_handle* h) {
  if (!h) return CDB_E_NULL;
  Comm* c = h->comm;
  if (!c) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int q = 0; q < c->world; ++q)
    if (q != c->rank && c->mail[q]) cudaIpcCloseMemHandle(c->mail[q]);
  if (c->local) cudaFree(c->local);
  delete c;
  h->comm = nullptr;
  return 0;
}

}  // extern "C"
