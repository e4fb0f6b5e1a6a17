// m3tsz_b200.hpp -- C++ host-side mirror of the reference's codec interface over
// the C ABI (include/m3tsz_b200.h).  Header-only.
//
// The reference is Go; with no Go toolchain in the build image the host side
// above the C ABI is C++ (this file) plus the Python mirror in
// m3_b200/encoding.py.  Names follow the reference interfaces
// (src/dbnode/encoding/types.go:39-91,180-203,342-345):
//   m3tsz::Encoder        <- encoding.Encoder        (m3tsz/encoder.go)
//   m3tsz::ReaderIterator <- encoding.ReaderIterator (m3tsz/iterator.go)
//   m3tsz::Decoder        <- encoding.Decoder        (m3tsz/decoder.go)
//   m3tsz::BatchCodec     <- the batch sites
// Encoder / ReaderIterator are RAII wrappers of the C streaming handles (m3tsz_encoder_*,
// m3tsz_iter_*): the bitstream is produced / consumed by one GPU launch when it is needed.
// There is no CPU codec behind this header.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "m3tsz_b200.h"

namespace m3tsz {

struct Error : std::runtime_error {
  int status;
  Error(int st, const std::string &what)
      : std::runtime_error(std::string(m3tsz_status_string(st)) + (what.empty() ? "" : ": " + what)),
        status(st) {}
};

struct Datapoint {  // ts.Datapoint, src/dbnode/ts/types.go:49-67
  int64_t timestamp_nanos;
  double value;
};

// encoding.Options subset + intOptimized (m3tsz.DefaultIntOptimizationEnabled = true)
struct Options {
  bool int_optimized = true;
  int default_time_unit = M3TSZ_UNIT_SECOND;
};

// One m3tsz_ctx per GPU; shared by the objects below (single-threaded use, like
// the reference's encoders / iterators).
class BatchCodec {
 public:
  explicit BatchCodec(int device = 0, Options o = Options()) {
    opts_.int_optimized = o.int_optimized ? 1 : 0;
    opts_.default_time_unit = o.default_time_unit;
    int rc = m3tsz_ctx_create(device, &ctx_);
    if (rc != M3TSZ_OK) throw Error(rc, "m3tsz_ctx_create");
  }
  ~BatchCodec() { m3tsz_ctx_destroy(ctx_); }
  BatchCodec(const BatchCodec &) = delete;
  BatchCodec &operator=(const BatchCodec &) = delete;

  // N x (Reset; Encode...; Discard).  ts/vals: [n_series][stride]; returns packed
  // streams + CSR offsets (n_series + 1) + per-series status.
  void EncodeBatch(const int64_t *ts, const double *vals, uint64_t n_series, uint64_t stride,
                   const uint32_t *n_points, const int64_t *starts, int unit,
                   std::vector<uint8_t> &packed, std::vector<uint64_t> &offsets,
                   std::vector<int32_t> &status, const uint8_t *units = nullptr,
                   const uint64_t *ann_series_off = nullptr,
                   const m3tsz_annotation_entry *ann_entries = nullptr,
                   const uint8_t *ann_bytes = nullptr, uint64_t ann_bytes_len = 0) {
    uint64_t extra = ann_series_off ? ann_bytes_len + 16 * ann_series_off[n_series] : 0;
    packed.resize(n_series * (m3tsz_encode_bound(stride) + extra) + 16);
    offsets.assign(n_series + 1, 0);
    status.assign(n_series, 0);
    std::vector<uint64_t> lens(n_series);
    int rc = m3tsz_encode_batch_host(ctx_, &opts_, ts, vals, n_series, stride, n_points, starts, unit,
                                     units, ann_series_off, ann_entries, ann_bytes, ann_bytes_len, 1,
                                     packed.data(), packed.size(), offsets.data(), lens.data(),
                                     status.data());
    if (rc != M3TSZ_OK) throw Error(rc, last_error());
    packed.resize(offsets[n_series]);
  }

  // N x (NewReaderIterator; for Next() { Current() }).  Outputs [n_series][max_points].
  void DecodeBatch(const uint8_t *streams, uint64_t streams_bytes, const uint64_t *offsets,
                   uint64_t n_series, uint64_t max_points, std::vector<int64_t> &ts,
                   std::vector<double> &vals, std::vector<uint32_t> &n_points,
                   std::vector<int32_t> &status, std::vector<uint8_t> *unit = nullptr,
                   std::vector<m3tsz_annotation_ref> *ann = nullptr) {
    ts.resize(n_series * max_points);
    vals.resize(n_series * max_points);
    n_points.assign(n_series, 0);
    status.assign(n_series, 0);
    if (unit) unit->assign(n_series, 0);
    if (ann) ann->assign(n_series, m3tsz_annotation_ref{0, 0, 0});
    int rc = m3tsz_decode_batch_host(ctx_, &opts_, streams, streams_bytes, offsets, n_series, ts.data(),
                                     vals.data(), max_points, n_points.data(), status.data(),
                                     unit ? unit->data() : nullptr, ann ? ann->data() : nullptr);
    if (rc != M3TSZ_OK) throw Error(rc, last_error());
  }

  std::string last_error() const { return m3tsz_last_cuda_error(ctx_); }
  m3tsz_ctx *ctx() { return ctx_; }
  const m3tsz_options &options() const { return opts_; }

 private:
  m3tsz_ctx *ctx_ = nullptr;
  m3tsz_options opts_{};
};

// encoding.Encoder (m3tsz/encoder.go:64-457): RAII over the C streaming handle; every method is
// one m3tsz_encoder_* call (the same entry points the cgo shim of INTEGRATION.md binds).
class Encoder {
 public:
  Encoder(BatchCodec &codec, int64_t start_nanos) {
    int rc = m3tsz_encoder_create(codec.ctx(), &codec.options(), start_nanos, &h_);
    if (rc != M3TSZ_OK) throw Error(rc, "m3tsz_encoder_create");
  }
  ~Encoder() { m3tsz_encoder_destroy(h_); }
  Encoder(const Encoder &) = delete;
  Encoder &operator=(const Encoder &) = delete;

  void Encode(Datapoint dp, int unit, const std::string &annotation = std::string()) {  // :90-110
    int rc = m3tsz_encoder_encode(h_, dp.timestamp_nanos, dp.value, unit,
                                  reinterpret_cast<const uint8_t *>(annotation.data()), annotation.size());
    if (rc == M3TSZ_ERR_DOD_OVERFLOW)
      throw Error(rc, "deltaOfDelta value " + std::to_string(m3tsz_encoder_failed_dod(h_)) +
                          (unit == M3TSZ_UNIT_SECOND ? " s" : " ms") + " overflows 32 bits");
    if (rc != M3TSZ_OK) throw Error(rc, "");
  }
  int NumEncoded() const { return (int)m3tsz_encoder_num_encoded(h_); }  // :299-302
  bool Empty() const { return m3tsz_encoder_empty(h_) != 0; }            // :330-332
  Datapoint LastEncoded() {  // :305-319, incl. the scaled-int / zero quirk
    Datapoint dp{0, 0.0};
    int rc = m3tsz_encoder_last_encoded(h_, &dp.timestamp_nanos, &dp.value);
    if (rc != M3TSZ_OK) throw Error(rc, "");
    return dp;
  }
  uint64_t LastAnnotationChecksum() const {  // :321-327
    uint64_t c = 0;
    int rc = m3tsz_encoder_last_annotation_checksum(h_, &c);
    if (rc != M3TSZ_OK) throw Error(rc, "");
    return c;
  }
  size_t Len() {  // :336-354
    uint64_t n = 0;
    int rc = m3tsz_encoder_len(h_, &n);
    if (rc != M3TSZ_OK) throw Error(rc, "");
    return (size_t)n;
  }
  // Stream(): head||tail bytes of the segment; empty vector == (nil, false)  (:282-297)
  const std::vector<uint8_t> &Stream() {
    bytes_.resize(Len());
    uint64_t n = 0;
    int rc = m3tsz_encoder_stream(h_, bytes_.data(), bytes_.size(), &n, &tail_len_);
    if (rc != M3TSZ_OK) throw Error(rc, "");
    bytes_.resize(n);
    return bytes_;
  }
  size_t TailLen() const { return (size_t)tail_len_; }  // ts.Segment tail of the last Stream()
  void Reset(int64_t start_nanos, uint64_t capacity = 0) {  // :262-279
    int rc = m3tsz_encoder_reset(h_, start_nanos, capacity);
    if (rc != M3TSZ_OK) throw Error(rc, "");
  }
  void Close() { m3tsz_encoder_close(h_); }  // :357-370
  std::vector<uint8_t> Discard() {            // :374-381
    std::vector<uint8_t> b = Stream();
    Close();
    return b;
  }
  std::vector<uint8_t> DiscardReset(int64_t start_nanos, uint64_t capacity = 0) {  // :385-392
    std::vector<uint8_t> b = Stream();
    Reset(start_nanos, capacity);
    return b;
  }

 private:
  m3tsz_encoder *h_ = nullptr;
  std::vector<uint8_t> bytes_;
  uint64_t tail_len_ = 0;
};

// encoding.ReaderIterator (m3tsz/iterator.go:67-278) over m3tsz_iter_*
class ReaderIterator {
 public:
  ReaderIterator(BatchCodec &codec, const uint8_t *data, size_t len) {
    int rc = m3tsz_iter_create(codec.ctx(), &codec.options(), &h_);
    if (rc != M3TSZ_OK) throw Error(rc, "m3tsz_iter_create");
    Reset(data, len);
  }
  ~ReaderIterator() { m3tsz_iter_destroy(h_); }
  ReaderIterator(ReaderIterator &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  ReaderIterator(const ReaderIterator &) = delete;
  ReaderIterator &operator=(const ReaderIterator &) = delete;

  void Reset(const uint8_t *data, size_t len) {  // :253-263
    int rc = m3tsz_iter_reset(h_, data, len);
    if (rc != M3TSZ_OK) throw Error(rc, "");
  }
  bool Next() { return m3tsz_iter_next(h_) != 0; }  // :81-106
  Datapoint Current() const {                       // :229-231
    Datapoint dp{0, 0.0};
    m3tsz_iter_current(h_, &dp.timestamp_nanos, &dp.value, nullptr, nullptr, nullptr);
    return dp;
  }
  int CurrentUnit() const {  // the unit in force at the current datapoint
    int32_t u = 0;
    m3tsz_iter_current(h_, nullptr, nullptr, &u, nullptr, nullptr);
    return u;
  }
  std::string CurrentAnnotation() const {  // annotation of the current datapoint ("" = none)
    const uint8_t *p = nullptr;
    uint64_t n = 0;
    m3tsz_iter_current(h_, nullptr, nullptr, nullptr, &p, &n);
    return n ? std::string(reinterpret_cast<const char *>(p), n) : std::string();
  }
  int Err() const { return m3tsz_iter_err(h_); }  // :234-236
  void Close() { m3tsz_iter_close(h_); }          // :267-278

 private:
  m3tsz_iter *h_ = nullptr;
};

// encoding.Decoder (m3tsz/decoder.go:27-44)
class Decoder {
 public:
  explicit Decoder(BatchCodec &codec) : codec_(codec) {}
  ReaderIterator Decode(const uint8_t *data, size_t len) { return ReaderIterator(codec_, data, len); }

 private:
  BatchCodec &codec_;
};

}  // namespace m3tsz
