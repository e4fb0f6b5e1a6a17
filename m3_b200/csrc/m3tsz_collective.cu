// m3tsz_collective.cu -- the one exchange step of the path (BASELINE config 5):
// a query that spans shards needs the DECODED blocks of every shard on every GPU.  Series
// shard trivially (the reference hash-partitions them, src/dbnode/sharding/shardset.go:157-173),
// so encode and decode run with no collective; this entry point decodes the local shard chunk
// by chunk and all-gathers chunk k-1 over NCCL / NVLink while chunk k is being decoded
// (two streams, no staging): every chunk is decoded into its final place and all-gathered in place,
// the gathered arrays are therefore CHUNK-major: [chunk][rank][chunk_series][point].
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the instance the process already
// loaded, e.g. torch's), so libm3tsz_b200.so has no link-time NCCL dependency and single-GPU
// users never touch it.
#include <dlfcn.h>

#include "m3tsz_ctx.h"

using namespace m3tsz;
using namespace m3tsz::host;

namespace {

// the handful of NCCL entry points used (nccl.h: ncclResult_t == int, ncclSuccess == 0,
// ncclInt8 == 0, ncclComm_t / ncclUniqueId opaque)
struct Uid {
  char internal[128];
};
struct Nccl {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, Uid /* ncclUniqueId by value: 128 bytes */, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*CommUserRank)(void *, int *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};

Nccl &nccl() {
  static Nccl n = [] {
    Nccl t;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      t.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (t.lib) break;
    }
    if (!t.lib) return t;
    t.GetUniqueId = (int (*)(void *))dlsym(t.lib, "ncclGetUniqueId");
    t.CommInitRank = (int (*)(void **, int, Uid, int))dlsym(t.lib, "ncclCommInitRank");
    t.CommDestroy = (int (*)(void *))dlsym(t.lib, "ncclCommDestroy");
    t.GroupStart = (int (*)())dlsym(t.lib, "ncclGroupStart");
    t.GroupEnd = (int (*)())dlsym(t.lib, "ncclGroupEnd");
    t.Broadcast = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(t.lib, "ncclBroadcast");
    t.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(t.lib, "ncclAllGather");
    t.CommUserRank = (int (*)(void *, int *))dlsym(t.lib, "ncclCommUserRank");
    t.GetErrorString = (const char *(*)(int))dlsym(t.lib, "ncclGetErrorString");
    t.ok = t.GetUniqueId && t.CommInitRank && t.CommDestroy && t.GroupStart && t.GroupEnd && t.Broadcast &&
           t.AllGather && t.CommUserRank;
    return t;
  }();
  return n;
}

int nccl_fail(m3tsz_ctx *ctx, int rc, const char *where) {
  Nccl &n = nccl();
  if (ctx)
    snprintf(ctx->last_error, sizeof(ctx->last_error), "%s: NCCL: %s", where,
             (n.GetErrorString ? n.GetErrorString(rc) : "error"));
  return M3TSZ_ERR_CUDA;
}

}  // namespace

extern "C" {

int m3tsz_nccl_unique_id(uint8_t *out128) {
  if (!out128) return M3TSZ_ERR_INVALID_ARG;
  Nccl &n = nccl();
  if (!n.ok) return M3TSZ_ERR_NO_DEVICE;
  Uid id;
  if (n.GetUniqueId(&id) != 0) return M3TSZ_ERR_CUDA;
  memcpy(out128, id.internal, 128);
  return M3TSZ_OK;
}

int m3tsz_nccl_comm_create(m3tsz_ctx *ctx, const uint8_t *unique_id128, int n_ranks, int rank, void **comm) {
  if (!ctx || !unique_id128 || !comm || n_ranks < 1 || rank < 0 || rank >= n_ranks) return M3TSZ_ERR_INVALID_ARG;
  Nccl &n = nccl();
  if (!n.ok) return M3TSZ_ERR_NO_DEVICE;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  Uid id;
  memcpy(id.internal, unique_id128, 128);
  int rc = n.CommInitRank(comm, n_ranks, id, rank);
  if (rc != 0) return nccl_fail(ctx, rc, "ncclCommInitRank");
  return M3TSZ_OK;
}

int m3tsz_nccl_comm_destroy(void *comm) {
  Nccl &n = nccl();
  if (!n.ok || !comm) return M3TSZ_ERR_INVALID_ARG;
  return n.CommDestroy(comm) == 0 ? M3TSZ_OK : M3TSZ_ERR_CUDA;
}

int m3tsz_allgather_decoded(m3tsz_ctx *ctx, const m3tsz_options *opts, void *nccl_comm, int n_ranks,
                            const uint8_t *d_streams, uint64_t streams_bytes, const uint64_t *d_offsets,
                            const uint64_t *d_lengths, uint64_t n_series, uint64_t gather_series,
                            uint64_t max_points, uint64_t chunk_series, int64_t *d_ts_all, double *d_val_all,
                            uint32_t *d_n_points_all, int32_t *d_status_all, void *stream) {
  if (!ctx || !valid_opts(opts) || !nccl_comm || n_ranks < 1) return M3TSZ_ERR_INVALID_ARG;
  if (gather_series == 0) return M3TSZ_OK;
  if (!d_streams || !d_offsets || !d_ts_all || !d_val_all || !d_n_points_all || !d_status_all || max_points == 0 ||
      gather_series > n_series)
    return M3TSZ_ERR_INVALID_ARG;
  Nccl &n = nccl();
  if (!n.ok) return M3TSZ_ERR_NO_DEVICE;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  if (chunk_series == 0) chunk_series = 32768;
  if (chunk_series > gather_series) chunk_series = gather_series;
  if (gather_series % chunk_series != 0) return M3TSZ_ERR_INVALID_ARG;  // whole chunks (see the layout)
  cudaStream_t user = (cudaStream_t)stream, dec = ctx->stream, com = ctx->stream2;
  int rank = -1;
  int rc = n.CommUserRank(nccl_comm, &rank);
  if (rc != 0 || rank < 0 || rank >= n_ranks) return nccl_fail(ctx, rc ? rc : 5, "ncclCommUserRank");
  cudaEvent_t ev_dec[2], ev_start;
  for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&ev_dec[i], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
  // everything queued on the caller's stream so far happens before the pipeline
  CK(cudaEventRecord(ev_start, user));
  CK(cudaStreamWaitEvent(dec, ev_start, 0));
  CK(cudaStreamWaitEvent(com, ev_start, 0));
  // No staging: chunk k is decoded straight into this rank's block of chunk k of the gathered arrays
  // ([chunk][rank][chunk_series][point]) and all-gathered IN PLACE (sendbuff = recvbuff + rank * count), so
  // the decodes run ahead on one stream while the gathers follow on the other.  n_points / status are
  // [rank][gather_series]: decoded into this rank's row, gathered once at the end.
  const uint64_t n_chunks = gather_series / chunk_series, C = chunk_series;
  uint32_t *my_n = d_n_points_all + (uint64_t)rank * gather_series;
  int32_t *my_st = d_status_all + (uint64_t)rank * gather_series;
  int status = M3TSZ_OK;
  for (uint64_t k = 0; k < n_chunks && status == M3TSZ_OK; k++) {
    const int slot = (int)(k & 1);
    const uint64_t s0 = k * C;
    const uint64_t blk = (k * (uint64_t)n_ranks) * C * max_points;       // chunk k of the gathered arrays
    const uint64_t mine = blk + (uint64_t)rank * C * max_points;          // this rank's block in it
    m3tsz_decode_extras ex;
    memset(&ex, 0, sizeof(ex));
    ex.d_lengths = d_lengths ? d_lengths + s0 : nullptr;
    status = m3tsz_decode_batch_ex(ctx, opts, d_streams, streams_bytes, d_offsets + s0, C, d_ts_all + mine,
                                   d_val_all + mine, max_points, my_n + s0, my_st + s0, nullptr, nullptr,
                                   d_lengths ? &ex : nullptr, dec);
    if (status != M3TSZ_OK) break;
    CK(cudaEventRecord(ev_dec[slot], dec));
    CK(cudaStreamWaitEvent(com, ev_dec[slot], 0));
    // (grouped ncclBroadcasts with strided placement -- a [rank][series][point] result -- measured
    // 316 GB/s per GPU at N = 2, staged ncclAllGather 388: the layout follows the collective)
    int nrc = n.GroupStart();
    if (nrc == 0) nrc = n.AllGather(d_ts_all + mine, d_ts_all + blk, C * max_points * 8, 0 /* ncclInt8 */, nccl_comm, com);
    if (nrc == 0) nrc = n.AllGather(d_val_all + mine, d_val_all + blk, C * max_points * 8, 0, nccl_comm, com);
    const int erc = n.GroupEnd();
    if (nrc != 0 || erc != 0) {
      status = nccl_fail(ctx, nrc ? nrc : erc, "ncclAllGather (decoded blocks)");
      break;
    }
    ctx->launches++;
  }
  if (status == M3TSZ_OK) {
    int nrc = n.GroupStart();
    if (nrc == 0) nrc = n.AllGather(my_n, d_n_points_all, gather_series * 4, 0, nccl_comm, com);
    if (nrc == 0) nrc = n.AllGather(my_st, d_status_all, gather_series * 4, 0, nccl_comm, com);
    const int erc = n.GroupEnd();
    if (nrc != 0 || erc != 0) status = nccl_fail(ctx, nrc ? nrc : erc, "ncclAllGather (n_points / status)");
  }
  // the caller's stream continues after the whole pipeline
  cudaEventRecord(ev_start, com);
  cudaStreamWaitEvent(user, ev_start, 0);
  cudaEventRecord(ev_start, dec);
  cudaStreamWaitEvent(user, ev_start, 0);
  for (int i = 0; i < 2; i++) cudaEventDestroy(ev_dec[i]);
  cudaEventDestroy(ev_start);
  return status;
}

}  // extern "C"
