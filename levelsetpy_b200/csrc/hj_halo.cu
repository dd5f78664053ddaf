// hj_halo.cu -- slab halos over NVLink peer memory (C-ABI: hj_halo_export / attach / push / wait / detach).
//
// a grid decomposed into slabs along dim 0 exchanges a 3-plane halo with each neighbour per RK stage
// (the reference has no multi-GPU path; the exchange replaces the ghost planes addGhostExtrapolate / addGhostPeriodic
// would have read from the same array, add_ghost_periodic.py:78-87).  Here a rank PUSHES its edge planes straight
// into the neighbour's stored halo planes:
//   * every context exports its three RK buffers and a small flag array as CUDA IPC handles (plain bytes the host
//     moves with whatever it has -- torch.distributed, MPI, a pipe); the neighbour maps them (cudaIpcOpenMemHandle;
//     contexts of the same process use the raw pointers);
//   * hj_halo_push: behind an event on the producing stream, one cudaMemcpyAsync per neighbour on a dedicated
//     high-priority stream -- the copy engines move the planes over NVLink, no SM runs a copy kernel, so the transfer
//     does not contend with the stage kernel running under it -- followed by a release store of the push counter into
//     the neighbour's flag;
//   * hj_halo_wait: the consuming stream waits (one spinning thread, acquire loads of its own flag) until both
//     neighbours' planes of that buffer have landed, and until its own outbound copies have left.
// Counters only grow: a buffer is pushed once and awaited once per step by every rank, so "push k has landed" is
// flag >= k.  A push into a neighbour's halo of buffer b cannot overtake the neighbour's last read of it: between two
// pushes of the same buffer lie two other stages, each of which completes a push/wait handshake with that neighbour.
#include <unistd.h>

#include <cstring>

#include "hj_ctx.h"

static const int HALO_LANES = 4;                 // copy streams (-> copy engines) per neighbour
static const size_t HALO_LANE_MIN = 32u << 20;   // a piece is at least this many bytes

struct HjHaloPeer {
  bool present = false, mapped = false;
  double* buf[3] = {};                    // the neighbour's RK buffers (base, halo planes included)
  unsigned long long* flags = nullptr;    // the neighbour's arrival counters [2 sides][3 buffers]
  long long n0 = 0;                       // the neighbour's slab height (locates its upper halo)
  cudaStream_t cs[HALO_LANES] = {};       // copy streams towards this neighbour (a large face is cut into pieces so
  cudaEvent_t ev_lane[HALO_LANES] = {};   // that several copy engines work on it); lane 0 also carries the flag
  cudaEvent_t ev_sent = nullptr;          // my last push towards it has left my buffers
};

struct HjHalo {
  unsigned long long* flags = nullptr;    // my arrival counters: [side of MY halo: 0 lower, 1 upper][buffer]
  HjHaloPeer peer[2];                     // 0 = lower neighbour, 1 = upper neighbour
  unsigned long long pushed[2][3] = {}, waited[3] = {};   // pushes towards [lower, upper] neighbour / waits, per buffer
  cudaEvent_t ev_ready = nullptr;         // "the planes to push are written" on the producing stream
  int fused = 0;                          // pass 2 of the split path stores its edge planes into the lower (bit 0) /
                                          // upper (bit 1) neighbour itself
};

namespace {

__global__ void k_flag_set(unsigned long long* flag, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(v) : "memory");
}

__global__ void k_flag_wait(const unsigned long long* flag, unsigned long long v) {
  unsigned long long cur;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(flag) : "memory");
    if (cur >= v) break;
    __nanosleep(200);
  }
}

struct Desc {                             // layout of the HJ_HALO_DESC_BYTES blob (same build on both sides)
  cudaIpcMemHandle_t buf[3];
  cudaIpcMemHandle_t flags;
  unsigned long long buf_ptr[3];
  unsigned long long flags_ptr;
  long long plane, n0;
  int device, pid;
};
static_assert(sizeof(Desc) <= HJ_HALO_DESC_BYTES, "halo descriptor does not fit its blob");

int ensure_halo(hj_ctx* c) {
  if (c->halo) return HJ_OK;
  HjHalo* h = new HjHalo();
  cudaError_t e = cudaMalloc(&h->flags, 6 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->flags, 0, 6 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();     // the zeroed counters are visible before anyone maps them
  if (e != cudaSuccess) {
    cudaFree(h->flags);
    delete h;
    return hj_fail(HJ_ERR_CUDA, "halo transport: allocation failed: %s", cudaGetErrorString(e));
  }
  c->halo = h;
  return HJ_OK;
}

void close_peer(HjHaloPeer& p) {
  if (p.mapped) {
    for (int b = 0; b < 3; ++b) cudaIpcCloseMemHandle(p.buf[b]);
    cudaIpcCloseMemHandle(p.flags);
  }
  if (p.ev_sent) cudaEventDestroy(p.ev_sent);
  for (int l = 0; l < HALO_LANES; ++l) {
    if (p.ev_lane[l]) cudaEventDestroy(p.ev_lane[l]);
    if (p.cs[l]) cudaStreamDestroy(p.cs[l]);
  }
  p = HjHaloPeer();
  (void)cudaGetLastError();
}

}  // namespace

void hj_halo_destroy(hj_ctx* c) {
  if (!c || !c->halo) return;
  // a wrap-around pair of two ranks maps the same neighbour on both sides: the mapping belongs to side 0
  if (c->halo->peer[1].mapped && c->halo->peer[0].mapped && c->halo->peer[1].flags == c->halo->peer[0].flags)
    c->halo->peer[1].mapped = false;
  for (int s = 0; s < 2; ++s) close_peer(c->halo->peer[s]);
  if (c->halo->ev_ready) cudaEventDestroy(c->halo->ev_ready);
  cudaFree(c->halo->flags);
  delete c->halo;
  c->halo = nullptr;
}

// peer addresses for the fused push of pass 2 (KStage::push_lo / push_hi): base such that base + (offset of a node of my
// 3 lowest / highest planes, relative to my first interior node) is that node's place in the neighbour's upper / lower
// halo of RK buffer `out_buf`
void hj_halo_fused_targets(hj_ctx* c, int out_buf, double** lo, double** hi) {
  *lo = *hi = nullptr;
  if (!c->halo || !c->halo->fused) return;
  const HjHaloPeer& pl = c->halo->peer[0];
  const HjHaloPeer& ph = c->halo->peer[1];
  if (pl.present && (c->halo->fused & 1)) *lo = pl.buf[out_buf] + (pl.n0 + HJ_GHOST) * c->plane;
  if (ph.present && (c->halo->fused & 2)) *hi = ph.buf[out_buf] - (long long)(c->gp.N[0] - HJ_GHOST) * c->plane;
}

extern "C" {

int hj_halo_set_fused(hj_ctx* c, int sides) {
  if (!c) return hj_fail(HJ_ERR_INVALID, "hj_halo_set_fused: null ctx");
  if (!c->halo) return hj_fail(HJ_ERR_STATE, "hj_halo_set_fused: no neighbour attached (hj_halo_attach)");
  if (sides < 0 || sides > 3) return hj_fail(HJ_ERR_INVALID, "hj_halo_set_fused: sides is a bit mask (1 lower, 2 upper)");
  c->halo->fused = sides;   // a plane in BOTH neighbours' halos (thin slabs) is simply stored twice
  return HJ_OK;
}

// after a pass-2 launch with fused pushes: bump the arrival counters of `buf` at the neighbours in `sides`, ordered
// behind `stream`
int hj_halo_signal(hj_ctx* c, void* stream, int buf, int sides) {
  if (!c || buf < 0 || buf > 2) return hj_fail(HJ_ERR_INVALID, "hj_halo_signal: bad argument");
  if (!c->halo) return hj_fail(HJ_ERR_STATE, "hj_halo_signal: no neighbour attached (hj_halo_attach)");
  HJ_CK(cudaSetDevice(c->device));
  HjHalo* h = c->halo;
  for (int side = 0; side < 2; ++side) {
    HjHaloPeer& p = h->peer[side];
    if (!p.present || !(sides & (1 << side))) continue;
    k_flag_set<<<1, 1, 0, (cudaStream_t)stream>>>(p.flags + (side == 1 ? 0 : 3) + buf, ++h->pushed[side][buf]);
    HJ_CK(cudaGetLastError());
  }
  return HJ_OK;
}

int hj_halo_export(hj_ctx* c, void* desc_out) {
  if (!c || !desc_out) return hj_fail(HJ_ERR_INVALID, "hj_halo_export: null argument");
  if (!c->halo0) return hj_fail(HJ_ERR_STATE, "hj_halo_export: context has no stored halo planes (dim 0 is not HJ_BC_HALO)");
  HJ_CK(cudaSetDevice(c->device));
  double* p0 = nullptr;
  int r = hj_state_ptr(c, 1, &p0);                       // allocates the RK buffers on first use
  if (r) return r;
  if ((r = ensure_halo(c))) return r;
  Desc d;
  std::memset(&d, 0, sizeof d);
  for (int b = 0; b < 3; ++b) {
    HJ_CK(cudaIpcGetMemHandle(&d.buf[b], c->buf[b]));
    d.buf_ptr[b] = (unsigned long long)c->buf[b];
  }
  HJ_CK(cudaIpcGetMemHandle(&d.flags, c->halo->flags));
  d.flags_ptr = (unsigned long long)c->halo->flags;
  d.plane = c->plane;
  d.n0 = c->gp.N[0];
  d.device = c->device;
  d.pid = (int)getpid();
  std::memset(desc_out, 0, HJ_HALO_DESC_BYTES);
  std::memcpy(desc_out, &d, sizeof d);
  return HJ_OK;
}

int hj_halo_attach(hj_ctx* c, const void* lower_desc, const void* upper_desc) {
  if (!c) return hj_fail(HJ_ERR_INVALID, "hj_halo_attach: null ctx");
  if (!c->halo0) return hj_fail(HJ_ERR_STATE, "hj_halo_attach: context has no stored halo planes");
  HJ_CK(cudaSetDevice(c->device));
  int r = ensure_halo(c);
  if (r) return r;
  HjHalo* h = c->halo;
  const void* descs[2] = {lower_desc, upper_desc};
  int prio_lo = 0, prio_hi = 0;
  HJ_CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));   // prio_hi = numerically lowest = greatest priority
  for (int s = 0; s < 2; ++s) {
    close_peer(h->peer[s]);
    if (!descs[s]) continue;
    Desc d;
    std::memcpy(&d, descs[s], sizeof d);
    if (d.plane != c->plane) return hj_fail(HJ_ERR_INVALID, "hj_halo_attach: neighbour's plane size differs (%lld vs %lld)", d.plane, c->plane);
    HjHaloPeer& p = h->peer[s];
    p.n0 = d.n0;
    if (d.pid == (int)getpid()) {                       // same process: the neighbour's pointers are valid here
      if (d.device != c->device) {
        int can = 0;
        HJ_CK(cudaDeviceCanAccessPeer(&can, c->device, d.device));
        if (!can) return hj_fail(HJ_ERR_UNSUPPORTED, "hj_halo_attach: no peer access from device %d to device %d", c->device, d.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(d.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) HJ_CK(e);
        (void)cudaGetLastError();
      }
      for (int b = 0; b < 3; ++b) p.buf[b] = (double*)d.buf_ptr[b];
      p.flags = (unsigned long long*)d.flags_ptr;
    } else if (s == 1 && h->peer[0].mapped && descs[0] && !std::memcmp(descs[0], descs[1], sizeof d)) {
      for (int b = 0; b < 3; ++b) p.buf[b] = h->peer[0].buf[b];   // two-rank ring: one neighbour on both sides
      p.flags = h->peer[0].flags;
      p.mapped = true;
    } else {
      for (int b = 0; b < 3; ++b) {
        void* q = nullptr;
        HJ_CK(cudaIpcOpenMemHandle(&q, d.buf[b], cudaIpcMemLazyEnablePeerAccess));
        p.buf[b] = (double*)q;
      }
      void* q = nullptr;
      HJ_CK(cudaIpcOpenMemHandle(&q, d.flags, cudaIpcMemLazyEnablePeerAccess));
      p.flags = (unsigned long long*)q;
      p.mapped = true;
    }
    for (int l = 0; l < HALO_LANES; ++l) {
      HJ_CK(cudaStreamCreateWithPriority(&p.cs[l], cudaStreamNonBlocking, prio_hi));
      HJ_CK(cudaEventCreateWithFlags(&p.ev_lane[l], cudaEventDisableTiming));
    }
    HJ_CK(cudaEventCreateWithFlags(&p.ev_sent, cudaEventDisableTiming));
    p.present = true;
  }
  return HJ_OK;
}

int hj_halo_detach(hj_ctx* c) {
  if (!c) return hj_fail(HJ_ERR_INVALID, "hj_halo_detach: null ctx");
  if (c->halo) {
    HJ_CK(cudaSetDevice(c->device));
    HJ_CK(cudaDeviceSynchronize());
    hj_halo_destroy(c);
  }
  return HJ_OK;
}

int hj_halo_attached(const hj_ctx* c) {
  if (!c || !c->halo) return 0;
  return (c->halo->peer[0].present ? 1 : 0) | (c->halo->peer[1].present ? 2 : 0);
}

// columns [col_begin, col_end) of every row of `row_len` elements of the three edge planes (row_len must divide the
// plane; 0, 0, 0 = the whole planes in one contiguous copy)
int hj_halo_push(hj_ctx* c, void* stream, int buf, int sides, int64_t col_begin, int64_t col_end, int64_t row_len) {
  if (!c || buf < 0 || buf > 2) return hj_fail(HJ_ERR_INVALID, "hj_halo_push: bad argument");
  if (!c->halo) return hj_fail(HJ_ERR_STATE, "hj_halo_push: no neighbour attached (hj_halo_attach)");
  const bool whole = row_len == 0;
  if (!whole && (row_len < 0 || c->plane % row_len || col_begin < 0 || col_end > row_len || col_begin >= col_end))
    return hj_fail(HJ_ERR_INVALID, "hj_halo_push: need 0 <= col_begin < col_end <= row_len, row_len dividing the plane");
  HJ_CK(cudaSetDevice(c->device));
  HjHalo* h = c->halo;
  cudaStream_t s = (cudaStream_t)stream;
  const long long n0 = c->gp.N[0], plane = c->plane;
  HJ_CK(cudaEventRecord(h->ev_ready, s));
  for (int side = 0; side < 2; ++side) {
    HjHaloPeer& p = h->peer[side];
    if (!p.present || !(sides & (1 << side))) continue;
    const unsigned long long seq = ++h->pushed[side][buf];
    // to the upper neighbour: my top 3 interior planes -> its lower halo; to the lower one: my bottom 3 -> its upper halo
    const double* src = c->buf[buf] + (side == 1 ? n0 : (long long)HJ_GHOST) * plane;
    double* dst = p.buf[buf] + (side == 1 ? 0 : (p.n0 + HJ_GHOST) * plane);
    unsigned long long* flag = p.flags + (side == 1 ? 0 : 3) + buf;
    // rows of the face: the whole 3 planes as one row, or the selected columns of every row_len-element row
    const size_t nrows = whole ? 1 : (size_t)(HJ_GHOST * plane / row_len);
    const size_t width = whole ? (size_t)HJ_GHOST * plane : (size_t)(col_end - col_begin);
    const size_t pitch = whole ? width : (size_t)row_len;
    const size_t total = nrows * width * sizeof(double);
    int lanes = (int)(total / HALO_LANE_MIN);
    lanes = lanes < 1 ? 1 : (lanes > HALO_LANES ? HALO_LANES : lanes);
    for (int l = 0; l < lanes; ++l) {
      HJ_CK(cudaStreamWaitEvent(p.cs[l], h->ev_ready, 0));
      if (whole) {
        const size_t a = width * l / lanes, b = width * (l + 1) / lanes;
        HJ_CK(cudaMemcpyAsync(dst + a, src + a, (b - a) * sizeof(double), cudaMemcpyDeviceToDevice, p.cs[l]));
      } else {
        const size_t a = nrows * l / lanes, b = nrows * (l + 1) / lanes;
        HJ_CK(cudaMemcpy2DAsync(dst + a * pitch + col_begin, pitch * sizeof(double), src + a * pitch + col_begin,
                                pitch * sizeof(double), width * sizeof(double), b - a, cudaMemcpyDeviceToDevice, p.cs[l]));
      }
      if (l) {
        HJ_CK(cudaEventRecord(p.ev_lane[l], p.cs[l]));
        HJ_CK(cudaStreamWaitEvent(p.cs[0], p.ev_lane[l], 0));
      }
    }
    k_flag_set<<<1, 1, 0, p.cs[0]>>>(flag, seq);          // behind every piece
    HJ_CK(cudaGetLastError());
    HJ_CK(cudaEventRecord(p.ev_sent, p.cs[0]));
  }
  return HJ_OK;
}

// `npush`: how many pushes of this buffer (per neighbour) to wait for -- 1 for whole planes, the number of column
// chunks when the neighbours push in pieces
int hj_halo_wait(hj_ctx* c, void* stream, int buf, int npush) {
  if (!c || buf < 0 || buf > 2 || npush < 1) return hj_fail(HJ_ERR_INVALID, "hj_halo_wait: bad argument");
  if (!c->halo) return hj_fail(HJ_ERR_STATE, "hj_halo_wait: no neighbour attached (hj_halo_attach)");
  HJ_CK(cudaSetDevice(c->device));
  HjHalo* h = c->halo;
  cudaStream_t s = (cudaStream_t)stream;
  h->waited[buf] += (unsigned long long)npush;
  for (int side = 0; side < 2; ++side) {
    if (!h->peer[side].present) continue;
    k_flag_wait<<<1, 1, 0, s>>>(h->flags + side * 3 + buf, h->waited[buf]);
    HJ_CK(cudaGetLastError());
    HJ_CK(cudaStreamWaitEvent(s, h->peer[side].ev_sent, 0));    // my own edge planes may be overwritten from here on
  }
  return HJ_OK;
}

}  // extern "C"
