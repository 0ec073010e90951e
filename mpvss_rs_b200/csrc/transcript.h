// Host side of the Fiat-Shamir transcript (dleq.rs:58-61, 87-99; participant.rs:238-252, 438-454).
//
// The device writes every participant's four framed elements  F(X) F(Y) F(a1) F(a2),
// F(e) = len_u64_be || bytes(e), into one fixed-size ROW per participant (frame kernels in modp.cu /
// ec.cu): a frame occupies 8 + EB bytes, the bytes left-aligned (ModpGroup elements are minimal-length
// big-endian, modp.rs:150-152, so a frame can be shorter than its slot; the curves' encodings have
// fixed length).  The host therefore hashes device-produced bytes as they are: one SHA-256 pass over
// the rows in `publickeys` order, no per-element byte reversal or length scan on the CPU.
//
// With N ranks (one GPU each) participant i lives on rank i % N at local row i / N; the rows of all ranks
// arrive through one all-gather as [rank][local row], and the hash walks them in participant order.
// The copy to the host is cut into chunks that are hashed while the next chunk is still in flight.
#pragma once
#include <stddef.h>
#include <algorithm>
#include <stdint.h>
#include "ctx.h"
#include "sha2.h"
#include "hash_launch.h"

namespace transcript {

struct Geom {
  size_t eb;       // element bytes at the boundary
  bool minimal;    // ModpGroup: frames carry minimal-length big-endian bytes
  size_t frame() const { return 8 + eb; }
  size_t row() const { return 4 * frame(); }
};

// rows per rank of a box of n_total participants dealt round robin to nranks
inline size_t rows_per_rank(size_t n_total, int nranks) { return (n_total + (size_t)nranks - 1) / (size_t)nranks; }
// participants owned by `rank`
inline size_t local_count(size_t n_total, int nranks, int rank) {
  return n_total > (size_t)rank ? (n_total - (size_t)rank + (size_t)nranks - 1) / (size_t)nranks : 0;
}

inline void hash_row(sha2::Sha256& h, const uint8_t* row, const Geom& g) {
  if (!g.minimal) {
    h.update(row, g.row());
    return;
  }
  const size_t f = g.frame();
  size_t len[4];
  bool full = true;
  for (int k = 0; k < 4; ++k) {
    len[k] = ((size_t)row[k * f + 6] << 8) | row[k * f + 7];
    full = full && len[k] == g.eb;
  }
  if (full) {
    h.update(row, g.row());
    return;
  }
  for (int k = 0; k < 4; ++k) h.update(row + k * f, 8 + len[k]);
}

// hash participants [i0, i1) of a gathered buffer laid out [rank][rows_per_rank][row]
inline void hash_range(sha2::Sha256& h, const uint8_t* gathered, size_t i0, size_t i1, int nranks, size_t rpr,
                       const Geom& g) {
  const size_t row = g.row();
  for (size_t i = i0; i < i1; ++i) hash_row(h, gathered + ((i % (size_t)nranks) * rpr + i / (size_t)nranks) * row, g);
}

// hash participants [i0, i1) of a buffer that is already in participant order: contiguous runs of rows whose
// frames are all full-length go to SHA-256 in one call (a 1056-byte update per row costs 20 % in partial-block
// handling), a row with a short frame is hashed frame by frame
inline void hash_ordered(sha2::Sha256& h, const uint8_t* rows, size_t i0, size_t i1, const Geom& g) {
  const size_t row = g.row(), f = g.frame();
  if (!g.minimal) {
    h.update(rows + i0 * row, (i1 - i0) * row);
    return;
  }
  size_t run = i0;
  for (size_t i = i0; i < i1; ++i) {
    const uint8_t* r = rows + i * row;
    bool full = true;
    for (int k = 0; k < 4; ++k) full = full && ((((size_t)r[k * f + 6] << 8) | r[k * f + 7]) == g.eb);
    if (!full) {
      if (i > run) h.update(rows + run * row, (i - run) * row);
      hash_row(h, r, g);
      run = i + 1;
    }
  }
  if (i1 > run) h.update(rows + run * row, (i1 - run) * row);
}

// The whole-box transcript hashed ON THE DEVICE ("device_hash" tunable): one thread walks the rows in
// participant order (shadev::box_hash_body) and 32 bytes come back instead of the rows.  SHA-256 is one sequential
// chain, so this is a single GPU thread against the host's SHA-NI unit: measured slower by two orders of magnitude
// (DESIGN.md section 5), kept as the checked alternative and never the default.  Also fetches the first byte
// of the first `nranks` rows, where a rank marks a slice that failed to decode.
inline int device_digest(mpvss_ctx* ctx, const uint8_t* dev_rows_ordered, size_t n_total, int nranks, const Geom& g,
                         uint8_t digest[32]) {
  DevBuf& dd = ctx->buf(20);
  MPVSS_CUDA(ctx, dd.ensure(32));
  shadev::BoxHashArgs A{dev_rows_ordered, (uint32_t)g.row(), (uint32_t)g.frame(), (uint32_t)n_total, dd.as<uint8_t>()};
  MPVSS_CUDA(ctx, shadev::launch_box_hash(A, ctx->stream));
  const size_t marks = std::min<size_t>((size_t)nranks, n_total);
  MPVSS_CUDA(ctx, ctx->h_frames.ensure(marks * g.row()));
  MPVSS_CUDA(ctx, cudaMemcpy2DAsync(ctx->h_frames.as<uint8_t>(), g.row(), dev_rows_ordered, g.row(), 1, marks,
                                    cudaMemcpyDeviceToHost, ctx->stream));
  MPVSS_CUDA(ctx, cudaMemcpyAsync(digest, dd.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
  MPVSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MPVSS_OK;
}

// Copy the transcript rows from the device (ctx->stream, after everything queued there) in chunks and hash them
// in participant order while later chunks are still copying.  `dev_rows` is in participant order
// (ordered = true: a single rank, or after reorder_rows) or in the all-gather layout [rank][local row].
inline int fetch_and_hash(mpvss_ctx* ctx, const uint8_t* dev_rows, size_t n_total, int nranks, const Geom& g,
                          sha2::Sha256& h, bool ordered = false) {
  const size_t rpr = rows_per_rank(n_total, nranks), row = g.row();
  if (nranks == 1) ordered = true;
  const size_t total_rows = ordered ? n_total : (size_t)nranks * rpr;
  MPVSS_CUDA(ctx, ctx->h_frames.ensure(total_rows * row));
  uint8_t* host = ctx->h_frames.as<uint8_t>();
  constexpr size_t NCH = sizeof(ctx->ev_chunk) / sizeof(ctx->ev_chunk[0]);
  if (ordered) {
    // about 2 MiB per chunk, at most NCH chunks
    const size_t rows_per_chunk = std::max<size_t>((2u << 20) / row, (n_total + NCH - 1) / NCH);
    const size_t nchunks = (n_total + rows_per_chunk - 1) / rows_per_chunk;
    for (size_t c = 0; c < nchunks; ++c) {
      const size_t i0 = c * rows_per_chunk, i1 = std::min(n_total, i0 + rows_per_chunk);
      MPVSS_CUDA(ctx, cudaMemcpyAsync(host + i0 * row, dev_rows + i0 * row, (i1 - i0) * row, cudaMemcpyDeviceToHost,
                                      ctx->stream));
      MPVSS_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[c], ctx->stream));
    }
    for (size_t c = 0; c < nchunks; ++c) {
      const size_t i0 = c * rows_per_chunk, i1 = std::min(n_total, i0 + rows_per_chunk);
      MPVSS_CUDA(ctx, cudaEventSynchronize(ctx->ev_chunk[c]));
      hash_ordered(h, host, i0, i1, g);
    }
    return MPVSS_OK;
  }
  // about 1 MiB per chunk and rank, at most NCH chunks
  size_t rows_per_chunk = std::max<size_t>((1u << 20) / row, (rpr + NCH - 1) / NCH);
  size_t nchunks = (rpr + rows_per_chunk - 1) / rows_per_chunk;
  for (size_t c = 0; c < nchunks; ++c) {
    const size_t j0 = c * rows_per_chunk, j1 = std::min(rpr, j0 + rows_per_chunk);
    for (int r = 0; r < nranks; ++r) {
      const size_t off = ((size_t)r * rpr + j0) * row;
      MPVSS_CUDA(ctx, cudaMemcpyAsync(host + off, dev_rows + off, (j1 - j0) * row, cudaMemcpyDeviceToHost, ctx->stream));
    }
    MPVSS_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[c], ctx->stream));
  }
  for (size_t c = 0; c < nchunks; ++c) {
    const size_t j0 = c * rows_per_chunk, j1 = std::min(rpr, j0 + rows_per_chunk);
    MPVSS_CUDA(ctx, cudaEventSynchronize(ctx->ev_chunk[c]));
    hash_range(h, host, std::min(n_total, j0 * (size_t)nranks), std::min(n_total, j1 * (size_t)nranks), nranks, rpr, g);
  }
  return MPVSS_OK;
}

}  // namespace transcript

// ---- communicator (comm.cu) -------------------------------------------------------------------
// all-gather `bytes` per rank from src into dst ([rank][bytes]) on ctx->stream; a plain device copy
// without a communicator
int comm_allgather(mpvss_ctx* ctx, const void* src, void* dst, size_t bytes);
void comm_release(mpvss_ctx* ctx);
// gathered rows [rank][rows_per_rank][row_bytes] -> participant order (row i = rank i % N, local row i / N), on the
// device at HBM speed, so that the host receives and hashes one contiguous transcript
int comm_reorder_rows(mpvss_ctx* ctx, const void* gathered, void* ordered, size_t n_total, size_t row_bytes);

namespace transcript {
// Bring this rank's n local rows of `kinds` row sets (device, widths[k] bytes per row) to the host in
// `publickeys` order for all n_total participants: plain copies without a communicator, else ONE all-gather of
// all row sets ([rank][kind][rows_per_rank][width]) and a scatter on the host.  host_all[k] may be null.
inline int gather_rows(mpvss_ctx* ctx, const void* const* dev_local, uint8_t* const* host_all, const size_t* widths,
                       int kinds, size_t n, size_t n_total) {
  if (ctx->nranks <= 1) {
    for (int k = 0; k < kinds; ++k)
      if (host_all[k])
        MPVSS_CUDA(ctx, cudaMemcpyAsync(host_all[k], dev_local[k], n * widths[k], cudaMemcpyDeviceToHost, ctx->stream));
    MPVSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPVSS_OK;
  }
  const size_t rpr = rows_per_rank(n_total, ctx->nranks), N = (size_t)ctx->nranks;
  size_t per_rank = 0, off[8] = {0};
  for (int k = 0; k < kinds; ++k) {
    off[k] = per_rank;
    per_rank += ((rpr * widths[k] + 15) / 16) * 16;
  }
  DevBuf& loc = ctx->buf(22);
  MPVSS_CUDA(ctx, loc.ensure(per_rank));
  MPVSS_CUDA(ctx, ctx->v_gather.ensure(N * per_rank));
  MPVSS_CUDA(ctx, cudaMemsetAsync(loc.p, 0, per_rank, ctx->stream));
  for (int k = 0; k < kinds; ++k)
    if (n)
      MPVSS_CUDA(ctx, cudaMemcpyAsync(loc.as<uint8_t>() + off[k], dev_local[k], n * widths[k], cudaMemcpyDeviceToDevice,
                                      ctx->stream));
  int st = comm_allgather(ctx, loc.p, ctx->v_gather.p, per_rank);
  if (st != MPVSS_OK) return st;
  MPVSS_CUDA(ctx, ctx->h_frames.ensure(N * per_rank));
  MPVSS_CUDA(ctx, cudaMemcpyAsync(ctx->h_frames.p, ctx->v_gather.p, N * per_rank, cudaMemcpyDeviceToHost, ctx->stream));
  MPVSS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const uint8_t* g = ctx->h_frames.as<uint8_t>();
  for (int k = 0; k < kinds; ++k) {
    if (!host_all[k]) continue;
    for (size_t i = 0; i < n_total; ++i)
      memcpy(host_all[k] + i * widths[k], g + (i % N) * per_rank + off[k] + (i / N) * widths[k], widths[k]);
  }
  return MPVSS_OK;
}
}  // namespace transcript
