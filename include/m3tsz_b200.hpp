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
// Per-datapoint methods are a buffered facade: the bitstream is produced /
// consumed by one GPU launch.  There is no CPU codec behind this header.
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

// encoding.Encoder facade (m3tsz/encoder.go:64-457)
class Encoder {
 public:
  Encoder(BatchCodec &codec, int64_t start_nanos) : codec_(codec) { Reset(start_nanos); }

  // Encode (encoder.go:90-110): validates what the reference validates eagerly
  // (closed encoder, unrecognised unit, s/ms delta-of-delta int32 overflow).
  void Encode(Datapoint dp, int unit, const std::string &annotation = std::string()) {
    if (closed_) throw Error(M3TSZ_ERR_ENCODER_CLOSED, "");
    const bool changed = unit >= 1 && unit <= 8 && unit != unit_;
    const int64_t delta = dp.timestamp_nanos - prev_time_;
    if (!changed) {
      if (unit < 1 || unit > 8) throw Error(M3TSZ_ERR_UNRECOGNIZED_UNIT, "");
      static const int64_t kNs[9] = {0, 1000000000LL, 1000000LL, 1000LL, 1LL, 60000000000LL,
                                     3600000000000LL, 86400000000000LL, 31536000000000000LL};
      const int64_t dod = (delta - prev_delta_) / kNs[unit];
      if (unit <= 2 && dod != (int64_t)(int32_t)dod)
        throw Error(M3TSZ_ERR_DOD_OVERFLOW, "deltaOfDelta value " + std::to_string(dod) +
                                                (unit == 1 ? " s" : " ms") + " overflows 32 bits");
    }
    if (!annotation.empty()) {
      m3tsz_annotation_entry e;
      e.dp_index = (uint32_t)ts_.size();
      e.length = (uint32_t)annotation.size();
      e.byte_offset = ann_bytes_.size();
      ann_entries_.push_back(e);
      ann_bytes_.insert(ann_bytes_.end(), annotation.begin(), annotation.end());
    }
    prev_time_ = dp.timestamp_nanos;
    if (changed) {
      unit_ = unit;
      prev_delta_ = 0;
    } else {
      prev_delta_ = delta;
    }
    ts_.push_back(dp.timestamp_nanos);
    vals_.push_back(dp.value);
    units_.push_back((uint8_t)unit);
    dirty_ = true;
  }
  int NumEncoded() const { return (int)ts_.size(); }   // :299-302
  bool Empty() const { return ts_.empty(); }           // :330-332
  Datapoint LastEncoded() const {                      // :305-319 (datapoint as written)
    if (ts_.empty()) throw Error(M3TSZ_ERR_NO_DATAPOINTS, "");
    return Datapoint{ts_.back(), vals_.back()};
  }
  // Stream(): head||tail bytes of the segment; empty vector == (nil, false)  (:282-297)
  const std::vector<uint8_t> &Stream() {
    if (dirty_) {
      std::vector<uint64_t> off;
      std::vector<int32_t> st;
      if (ts_.empty()) {
        bytes_.clear();
      } else {
        const uint64_t aoff[2] = {0, ann_entries_.size()};
        codec_.EncodeBatch(ts_.data(), vals_.data(), 1, ts_.size(), nullptr, &start_, M3TSZ_UNIT_SECOND,
                           bytes_, off, st, units_.data(), ann_entries_.empty() ? nullptr : aoff,
                           ann_entries_.data(), ann_bytes_.data(), ann_bytes_.size());
        if (st[0] != M3TSZ_OK) throw Error(st[0], "encode");
      }
      dirty_ = false;
    }
    return bytes_;
  }
  size_t Len() { return Stream().size(); }             // :336-354
  void Reset(int64_t start_nanos) {                    // :262-279
    start_ = start_nanos;
    ts_.clear();
    vals_.clear();
    units_.clear();
    ann_entries_.clear();
    ann_bytes_.clear();
    bytes_.clear();
    dirty_ = false;
    closed_ = false;
    prev_time_ = start_nanos;
    prev_delta_ = 0;
    const int du = codec_.options().default_time_unit;
    static const int64_t kNs[9] = {0, 1000000000LL, 1000000LL, 1000LL, 1LL, 60000000000LL,
                                   3600000000000LL, 86400000000000LL, 31536000000000000LL};
    unit_ = (du >= 1 && du <= 8 && start_nanos % kNs[du] == 0) ? du : 0;  // initialTimeUnit
  }
  void Close() { closed_ = true; }                     // :357-370
  std::vector<uint8_t> Discard() {                     // :374-381
    std::vector<uint8_t> b = Stream();
    Close();
    return b;
  }

 private:
  BatchCodec &codec_;
  int64_t start_ = 0, prev_time_ = 0, prev_delta_ = 0;
  int unit_ = 0;
  bool dirty_ = false, closed_ = false;
  std::vector<int64_t> ts_;
  std::vector<double> vals_;
  std::vector<uint8_t> units_, ann_bytes_, bytes_;
  std::vector<m3tsz_annotation_entry> ann_entries_;
};

// encoding.ReaderIterator facade (m3tsz/iterator.go:67-278)
class ReaderIterator {
 public:
  ReaderIterator(BatchCodec &codec, const uint8_t *data, size_t len) : codec_(codec) { Reset(data, len); }
  void Reset(const uint8_t *data, size_t len) {        // :253-263
    data_.assign(data, data + len);
    decoded_ = false;
    closed_ = false;
    i_ = -1;
    n_ = 0;
    err_ = 0;
  }
  bool Next() {                                        // :81-106
    if (closed_) return false;
    if (!decoded_) DecodeAll();
    if (i_ + 1 < n_) {
      i_++;
      return true;
    }
    i_ = n_;
    return false;
  }
  Datapoint Current() const { return Datapoint{ts_[i_], vals_[i_]}; }  // :229-231
  int CurrentUnit() const { return unit_; }
  int Err() const { return closed_ ? M3TSZ_ERR_ITER_CLOSED : ((i_ >= n_ - 1 || n_ == 0) ? err_ : 0); }
  void Close() { closed_ = true; }                     // :267-278

 private:
  void DecodeAll() {
    decoded_ = true;
    uint64_t cap = 2048;
    std::vector<uint32_t> n;
    std::vector<int32_t> st;
    std::vector<uint8_t> unit;
    const uint64_t off[2] = {0, data_.size()};
    for (;;) {
      codec_.DecodeBatch(data_.data(), data_.size(), off, 1, cap, ts_, vals_, n, st, &unit);
      if (st[0] == M3TSZ_ERR_CAPACITY) {
        cap = n[0];
        continue;
      }
      break;
    }
    n_ = (int64_t)(n[0] < cap ? n[0] : cap);
    err_ = st[0];
    unit_ = unit[0];
  }
  BatchCodec &codec_;
  std::vector<uint8_t> data_;
  std::vector<int64_t> ts_;
  std::vector<double> vals_;
  int64_t i_ = -1, n_ = 0;
  int err_ = 0, unit_ = 0;
  bool decoded_ = false, closed_ = false;
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
