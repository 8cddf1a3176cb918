// Transports of the slab decomposition: how the records packed by the engine's
// exchange kernels (mg.cuh) reach the neighbouring ranks.
//   NcclTransport   ncclSend / ncclRecv / ncclAllReduce on the context's stream
//                   (NVLink / NVSwitch between the GPUs of one box); NCCL is bound
//                   at run time (dlopen of libnccl.so.2 — the instance PyTorch has
//                   loaded when the host is Python, the system one otherwise).
//   HubTransport    ranks that live in ONE process (one host thread each): device
//                   to device copies ordered by CUDA events. Used to run several
//                   ranks on a single GPU (parity tests) with the very same
//                   pack / import kernels as the NCCL path.
// A transport moves DEVICE buffers; sizes are known on the host when a message is
// posted (the engine exchanges counts first where they vary).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <string>

namespace titgpu {

struct MgMsg {
  int peer;           // rank of the other side
  const void* send;   // device pointer (may be null when send_bytes == 0)
  size_t send_bytes;
  void* recv;         // device pointer (may be null when recv_bytes == 0)
  size_t recv_bytes;
};

struct MgTransport {
  virtual ~MgTransport() {}
  virtual int rank() const = 0;
  virtual int nranks() const = 0;
  // Exchange the messages (all of them in flight together). Stream-ordered with
  // respect to `stream`: data sent is what earlier work on `stream` produced, data
  // received is visible to later work on `stream`.
  virtual int sendrecv(cudaStream_t stream, const MgMsg* msgs, int nmsg, std::string& err) = 0;
  // Small host-side exchange of one 64-bit integer vector with each of the given peers
  // (counts). Blocks the host; `stream` is synchronised first by the caller.
  virtual int exchange_counts(cudaStream_t stream, const int* peers, int npeers, const long long* out, long long* in, int nvals, std::string& err) = 0;
  // In-place MIN / MAX over all ranks of two device words holding non-negative doubles
  // (which order like their bit patterns): the time-step scalars.
  virtual int allreduce_min_max(cudaStream_t stream, unsigned long long* d_min, unsigned long long* d_max, std::string& err) = 0;
  // Host-side sum over all ranks (diagnostics: particle-count conservation).
  virtual int allreduce_sum_host(cudaStream_t stream, long long* vals, int nvals, std::string& err) = 0;
};

// NCCL. `id128` = ncclUniqueId bytes produced by nccl_unique_id() on one rank.
int nccl_unique_id(void* id128, std::string& err);
MgTransport* make_nccl_transport(const void* id128, int rank, int nranks, int device, std::string& err);
// Adopt a communicator the host already owns (ncclComm_t); not destroyed by the transport.
MgTransport* adopt_nccl_comm(void* comm, int rank, int nranks, std::string& err);

// In-process hub shared by the ranks' transports.
struct MgHub;
MgHub* make_hub(int nranks);
void destroy_hub(MgHub*);
MgTransport* make_hub_transport(MgHub* hub, int rank, std::string& err);

}  // namespace titgpu
