// Transports of the slab decomposition (see mg_transport.h).
#include "mg_transport.h"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is bound with dlopen

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <vector>

namespace titgpu {

// ---------------------------------------------------------------------------
// NCCL, bound at run time.
// ---------------------------------------------------------------------------
namespace {

struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  std::string why;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // The instance already in the process (PyTorch's bundled NCCL) wins, so that both
    // sides of a host that also uses torch.distributed talk through one library.
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib)
      for (const char* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
      }
    if (!api.lib) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define TIT_NCCL_SYM(name)                                                   \
  api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.lib, "nccl" #name)); \
  if (!api.name) { api.why = "libnccl lacks nccl" #name; return; }
    TIT_NCCL_SYM(GetUniqueId)
    TIT_NCCL_SYM(CommInitRank)
    TIT_NCCL_SYM(CommDestroy)
    TIT_NCCL_SYM(Send)
    TIT_NCCL_SYM(Recv)
    TIT_NCCL_SYM(GroupStart)
    TIT_NCCL_SYM(GroupEnd)
    TIT_NCCL_SYM(AllReduce)
    TIT_NCCL_SYM(GetErrorString)
    TIT_NCCL_SYM(GetVersion)
#undef TIT_NCCL_SYM
  });
  return api;
}

#define TIT_NCCL_OK(expr)                                                              \
  do {                                                                                 \
    ncclResult_t r_ = (expr);                                                          \
    if (r_ != ncclSuccess) { err = std::string(#expr) + ": " + nccl_api().GetErrorString(r_); return 1; } \
  } while (0)
#define TIT_CU_OK(expr)                                                                \
  do {                                                                                 \
    cudaError_t e_ = (expr);                                                           \
    if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return 1; } \
  } while (0)

struct NcclTransport final : MgTransport {
  ncclComm_t comm = nullptr;
  bool owned = false;
  int rank_ = 0, nranks_ = 1;
  long long* d_cnt = nullptr;  // device scratch of exchange_counts: [out | in]
  long long* h_cnt = nullptr;  // pinned
  static constexpr int kCntCap = 64;

  ~NcclTransport() override {
    if (d_cnt) cudaFree(d_cnt);
    if (h_cnt) cudaFreeHost(h_cnt);
    if (comm && owned) nccl_api().CommDestroy(comm);
  }
  int rank() const override { return rank_; }
  int nranks() const override { return nranks_; }

  int scratch(std::string& err) {
    if (!d_cnt) TIT_CU_OK(cudaMalloc(&d_cnt, 2 * kCntCap * sizeof(long long)));
    if (!h_cnt) TIT_CU_OK(cudaMallocHost(&h_cnt, 2 * kCntCap * sizeof(long long)));
    return 0;
  }

  int sendrecv(cudaStream_t stream, const MgMsg* msgs, int nmsg, std::string& err) override {
    NcclApi& N = nccl_api();
    TIT_NCCL_OK(N.GroupStart());
    for (int i = 0; i < nmsg; ++i) {
      const MgMsg& m = msgs[i];
      if (m.send_bytes) TIT_NCCL_OK(N.Send(m.send, m.send_bytes, ncclChar, m.peer, comm, stream));
      if (m.recv_bytes) TIT_NCCL_OK(N.Recv(m.recv, m.recv_bytes, ncclChar, m.peer, comm, stream));
    }
    TIT_NCCL_OK(N.GroupEnd());
    return 0;
  }

  int exchange_counts(cudaStream_t stream, const int* peers, int npeers, const long long* out, long long* in, int nvals, std::string& err) override {
    if (npeers * nvals > kCntCap) { err = "exchange_counts: too many values"; return 1; }
    if (scratch(err)) return 1;
    NcclApi& N = nccl_api();
    const size_t tot = size_t(npeers) * nvals;
    std::memcpy(h_cnt, out, tot * sizeof(long long));
    TIT_CU_OK(cudaMemcpyAsync(d_cnt, h_cnt, tot * sizeof(long long), cudaMemcpyHostToDevice, stream));
    TIT_NCCL_OK(N.GroupStart());
    for (int p = 0; p < npeers; ++p) {
      TIT_NCCL_OK(N.Send(d_cnt + size_t(p) * nvals, nvals, ncclInt64, peers[p], comm, stream));
      TIT_NCCL_OK(N.Recv(d_cnt + kCntCap + size_t(p) * nvals, nvals, ncclInt64, peers[p], comm, stream));
    }
    TIT_NCCL_OK(N.GroupEnd());
    TIT_CU_OK(cudaMemcpyAsync(h_cnt + kCntCap, d_cnt + kCntCap, tot * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    TIT_CU_OK(cudaStreamSynchronize(stream));
    std::memcpy(in, h_cnt + kCntCap, tot * sizeof(long long));
    return 0;
  }

  int allreduce_min_max(cudaStream_t stream, unsigned long long* d_min, unsigned long long* d_max, std::string& err) override {
    NcclApi& N = nccl_api();
    TIT_NCCL_OK(N.GroupStart());
    TIT_NCCL_OK(N.AllReduce(d_min, d_min, 1, ncclUint64, ncclMin, comm, stream));
    TIT_NCCL_OK(N.AllReduce(d_max, d_max, 1, ncclUint64, ncclMax, comm, stream));
    TIT_NCCL_OK(N.GroupEnd());
    return 0;
  }

  int allreduce_sum_host(cudaStream_t stream, long long* vals, int nvals, std::string& err) override {
    if (nvals > kCntCap) { err = "allreduce_sum_host: too many values"; return 1; }
    if (scratch(err)) return 1;
    NcclApi& N = nccl_api();
    std::memcpy(h_cnt, vals, size_t(nvals) * sizeof(long long));
    TIT_CU_OK(cudaMemcpyAsync(d_cnt, h_cnt, size_t(nvals) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    TIT_NCCL_OK(N.AllReduce(d_cnt, d_cnt, nvals, ncclInt64, ncclSum, comm, stream));
    TIT_CU_OK(cudaMemcpyAsync(h_cnt, d_cnt, size_t(nvals) * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    TIT_CU_OK(cudaStreamSynchronize(stream));
    std::memcpy(vals, h_cnt, size_t(nvals) * sizeof(long long));
    return 0;
  }
};

}  // namespace

int nccl_unique_id(void* id128, std::string& err) {
  NcclApi& N = nccl_api();
  if (!N.GetUniqueId) { err = N.why.empty() ? "NCCL is not available" : N.why; return 1; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  TIT_NCCL_OK(N.GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof id);
  return 0;
}

MgTransport* make_nccl_transport(const void* id128, int rank, int nranks, int device, std::string& err) {
  NcclApi& N = nccl_api();
  if (!N.CommInitRank) { err = N.why.empty() ? "NCCL is not available" : N.why; return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { err = "cudaSetDevice failed"; return nullptr; }
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  NcclTransport* t = new NcclTransport();
  t->rank_ = rank; t->nranks_ = nranks; t->owned = true;
  const ncclResult_t r = N.CommInitRank(&t->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    err = std::string("ncclCommInitRank: ") + N.GetErrorString(r);
    t->comm = nullptr;
    delete t;
    return nullptr;
  }
  return t;
}

MgTransport* adopt_nccl_comm(void* comm, int rank, int nranks, std::string& err) {
  NcclApi& N = nccl_api();
  if (!N.Send) { err = N.why.empty() ? "NCCL is not available" : N.why; return nullptr; }
  if (!comm) { err = "null ncclComm_t"; return nullptr; }
  NcclTransport* t = new NcclTransport();
  t->comm = static_cast<ncclComm_t>(comm);
  t->rank_ = rank; t->nranks_ = nranks; t->owned = false;
  return t;
}

// ---------------------------------------------------------------------------
// In-process hub.
// ---------------------------------------------------------------------------
struct MgHub {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long gen = 0;
  bool failed = false;
  struct Slot {
    const MgMsg* msgs = nullptr;
    int nmsg = 0;
    cudaEvent_t ready = nullptr, done = nullptr;
    const int* peers = nullptr;
    const long long* out = nullptr;
    int npeers = 0, nvals = 0;
    unsigned long long mn = 0, mx = 0;
    std::vector<long long> sums;
    bool attached = false;
  };
  std::vector<Slot> slots;

  // Returns false when another rank failed or did not arrive within the time-out.
  bool barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (failed) return false;
    const unsigned long long g = gen;
    if (++arrived == n) {
      arrived = 0;
      ++gen;
      cv.notify_all();
      return true;
    }
    const bool ok = cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g || failed; });
    if (!ok) { failed = true; cv.notify_all(); }
    return ok && !failed;
  }
  void fail() {
    std::lock_guard<std::mutex> lk(m);
    failed = true;
    cv.notify_all();
  }
};

MgHub* make_hub(int nranks) {
  if (nranks < 1) return nullptr;
  MgHub* h = new MgHub();
  h->n = nranks;
  h->slots.resize(size_t(nranks));
  return h;
}
void destroy_hub(MgHub* h) { delete h; }

namespace {

struct HubTransport final : MgTransport {
  MgHub* hub = nullptr;
  int rank_ = 0;
  ~HubTransport() override {
    MgHub::Slot& s = hub->slots[size_t(rank_)];
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.done) cudaEventDestroy(s.done);
    s.ready = s.done = nullptr;
    s.attached = false;
  }
  int rank() const override { return rank_; }
  int nranks() const override { return hub->n; }

  int sendrecv(cudaStream_t stream, const MgMsg* msgs, int nmsg, std::string& err) override {
    MgHub::Slot& me = hub->slots[size_t(rank_)];
    if (cudaEventRecord(me.ready, stream) != cudaSuccess) { hub->fail(); err = "hub: cudaEventRecord failed"; return 1; }
    me.msgs = msgs; me.nmsg = nmsg;
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    int rc = 0;
    for (int i = 0; i < nmsg && !rc; ++i) {
      const MgMsg& m = msgs[i];
      if (!m.recv_bytes) continue;
      // the k-th message I exchange with this peer pairs with the k-th message the peer exchanges with me
      int k = 0;
      for (int j = 0; j < i; ++j) k += msgs[j].peer == m.peer;
      const MgHub::Slot& ps = hub->slots[size_t(m.peer)];
      const MgMsg* theirs = nullptr;
      for (int j = 0, kk = 0; j < ps.nmsg; ++j)
        if (ps.msgs[j].peer == rank_) { if (kk++ == k) { theirs = &ps.msgs[j]; break; } }
      if (!theirs || theirs->send_bytes != m.recv_bytes) { err = "hub: unmatched message sizes between ranks"; rc = 1; break; }
      if (cudaStreamWaitEvent(stream, ps.ready, 0) != cudaSuccess || cudaMemcpyAsync(m.recv, theirs->send, m.recv_bytes, cudaMemcpyDefault, stream) != cudaSuccess) {
        err = "hub: device copy failed"; rc = 1;
      }
    }
    if (rc) { hub->fail(); return 1; }
    if (cudaEventRecord(me.done, stream) != cudaSuccess) { hub->fail(); err = "hub: cudaEventRecord failed"; return 1; }
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    // my send buffers may be reused once the peers' copies are done
    for (int i = 0; i < nmsg; ++i)
      if (msgs[i].send_bytes && cudaStreamWaitEvent(stream, hub->slots[size_t(msgs[i].peer)].done, 0) != cudaSuccess) { hub->fail(); err = "hub: cudaStreamWaitEvent failed"; return 1; }
    // nobody may post the next round before everyone has read this one's tables
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    return 0;
  }

  int exchange_counts(cudaStream_t, const int* peers, int npeers, const long long* out, long long* in, int nvals, std::string& err) override {
    MgHub::Slot& me = hub->slots[size_t(rank_)];
    me.peers = peers; me.npeers = npeers; me.out = out; me.nvals = nvals;
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    int rc = 0;
    for (int p = 0; p < npeers && !rc; ++p) {
      const MgHub::Slot& ps = hub->slots[size_t(peers[p])];
      int q = -1;
      for (int j = 0; j < ps.npeers; ++j) if (ps.peers[j] == rank_) { q = j; break; }
      if (q < 0 || ps.nvals != nvals) { err = "hub: unmatched count exchange"; rc = 1; break; }
      std::memcpy(in + size_t(p) * nvals, ps.out + size_t(q) * nvals, size_t(nvals) * sizeof(long long));
    }
    if (rc) { hub->fail(); return 1; }
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    return 0;
  }

  int allreduce_min_max(cudaStream_t stream, unsigned long long* d_min, unsigned long long* d_max, std::string& err) override {
    MgHub::Slot& me = hub->slots[size_t(rank_)];
    if (cudaMemcpyAsync(&me.mn, d_min, 8, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaMemcpyAsync(&me.mx, d_max, 8, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) { hub->fail(); err = "hub: device copy failed"; return 1; }
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    unsigned long long mn = me.mn, mx = me.mx;
    for (const MgHub::Slot& s : hub->slots) { mn = s.mn < mn ? s.mn : mn; mx = s.mx > mx ? s.mx : mx; }
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    me.mn = mn; me.mx = mx;  // (own slot only: the peers have finished reading)
    if (cudaMemcpyAsync(d_min, &me.mn, 8, cudaMemcpyHostToDevice, stream) != cudaSuccess || cudaMemcpyAsync(d_max, &me.mx, 8, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) { hub->fail(); err = "hub: device copy failed"; return 1; }
    return 0;
  }

  int allreduce_sum_host(cudaStream_t, long long* vals, int nvals, std::string& err) override {
    MgHub::Slot& me = hub->slots[size_t(rank_)];
    me.sums.assign(vals, vals + nvals);
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    std::vector<long long> tot(size_t(nvals), 0);
    for (const MgHub::Slot& s : hub->slots)
      for (int i = 0; i < nvals && size_t(i) < s.sums.size(); ++i) tot[size_t(i)] += s.sums[size_t(i)];
    if (!hub->barrier()) { err = "hub: a peer rank failed or timed out"; return 1; }
    std::memcpy(vals, tot.data(), size_t(nvals) * sizeof(long long));
    return 0;
  }
};

}  // namespace

MgTransport* make_hub_transport(MgHub* hub, int rank, std::string& err) {
  if (!hub || rank < 0 || rank >= hub->n) { err = "hub: bad rank"; return nullptr; }
  MgHub::Slot& s = hub->slots[size_t(rank)];
  if (s.attached) { err = "hub: rank already attached"; return nullptr; }
  if (cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) {
    err = "hub: cudaEventCreate failed";
    return nullptr;
  }
  s.attached = true;
  HubTransport* t = new HubTransport();
  t->hub = hub; t->rank_ = rank;
  return t;
}

}  // namespace titgpu
