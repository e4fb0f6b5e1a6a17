// C++ facade test (include/m3tsz_b200.hpp): reproduces the reference's golden
// full streams (m3tsz/encoder_test.go:207-245,329-393 <-> iterator_test.go:181-222,315-385)
// through m3tsz::Encoder / m3tsz::ReaderIterator.  Exit 0 = pass, 3 = no GPU.
#include <cstdio>
#include <string>
#include <vector>

#include "m3tsz_b200.hpp"

static const int64_t SEC = 1000000000LL;

struct In {
  int64_t off;
  double v;
  int unit;
  std::string ann;
};

static int check(m3tsz::BatchCodec &codec, const std::vector<In> &in, const std::vector<uint8_t> &golden) {
  const int64_t start = 1427162400LL * SEC, t0 = 1427162462LL * SEC;
  m3tsz::Encoder enc(codec, start);
  if (!enc.Stream().empty() || enc.Len() != 0 || !enc.Empty()) return 10;
  for (const In &i : in) enc.Encode(m3tsz::Datapoint{t0 + i.off, i.v}, i.unit, i.ann);
  if (enc.Stream() != golden) return 11;
  if (enc.NumEncoded() != (int)in.size()) return 12;
  m3tsz::Decoder dec(codec);
  m3tsz::ReaderIterator it = dec.Decode(golden.data(), golden.size());
  size_t k = 0;
  while (it.Next()) {
    if (k >= in.size()) return 13;
    m3tsz::Datapoint dp = it.Current();
    if (dp.timestamp_nanos != t0 + in[k].off || dp.value != in[k].v) return 14;
    k++;
  }
  if (k != in.size() || it.Err() != 0) return 15;
  if (it.Next()) return 16;
  return 0;
}

int main() {
  try {
    m3tsz::Options o;
    o.int_optimized = false;
    m3tsz::BatchCodec codec(0, o);
    const int S = M3TSZ_UNIT_SECOND, MS = M3TSZ_UNIT_MILLISECOND;
    std::vector<In> a = {{0, 12, S, ""},          {60 * SEC, 12, S, ""},  {120 * SEC, 24, S, ""},
                         {-76 * SEC, 24, S, ""},  {-16 * SEC, 24, S, ""}, {2092 * SEC, 15, S, ""},
                         {4200 * SEC, 12, S, ""}};
    std::vector<uint8_t> ga = {0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0,  0x9f, 0x20, 0x14, 0x0,
                               0x0,  0x0,  0x0,  0x0,  0x0,  0x5f, 0x8c, 0xb0, 0x3a, 0x0,  0xe1, 0x0,
                               0x78, 0x0,  0x0,  0x40, 0x6,  0x58, 0x76, 0x8e, 0x0,  0x0};
    int rc = check(codec, a, ga);
    if (rc) {
      std::printf("FAIL no-annotation stream rc=%d\n", rc);
      return 1;
    }
    std::vector<In> b = {{0, 12, S, "\x0a"},
                         {60 * SEC, 12, S, ""},
                         {120 * SEC, 24, S, ""},
                         {-76 * SEC, 24, S, std::string("\x01\x02", 2)},
                         {-16 * SEC, 24, MS, ""},
                         {-15500 * 1000000LL, 15, MS, std::string("\x03\x04\x05", 3)},
                         {-14000 * 1000000LL, 12, S, ""}};
    std::vector<uint8_t> gb = {0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0,  0x80, 0x20, 0x1,  0x53, 0xe4,
                               0x2,  0x80, 0x0,  0x0,  0x0,  0x0,  0x0,  0xb,  0xf1, 0x96, 0x6,  0x0,  0x81,
                               0x0,  0x81, 0x68, 0x2,  0x1,  0x1,  0x0,  0x0,  0x0,  0x1d, 0xcd, 0x65, 0x0,
                               0x0,  0x20, 0x8,  0x20, 0x18, 0x20, 0x2f, 0xf,  0xa6, 0x58, 0x77, 0x0,  0x80,
                               0x40, 0x0,  0x0,  0x0,  0xe,  0xe6, 0xb2, 0x80, 0x23, 0x80, 0x0};
    rc = check(codec, b, gb);
    if (rc) {
      std::printf("FAIL annotation+time-unit stream rc=%d\n", rc);
      return 1;
    }
    // error behaviour: delta-of-delta overflow (encoder_test.go:581-651), closed encoder
    m3tsz::Encoder e(codec, 1427162400LL * SEC);
    e.Encode(m3tsz::Datapoint{1427162400LL * SEC, 1}, S);
    bool threw = false;
    try {
      e.Encode(m3tsz::Datapoint{1427162400LL * SEC + 1000LL * 25 * 24 * 3600 * SEC, 2}, S);
    } catch (const m3tsz::Error &err) {
      threw = err.status == M3TSZ_ERR_DOD_OVERFLOW &&
              std::string(err.what()).find("deltaOfDelta value 2160000000 s overflows 32 bits") != std::string::npos;
    }
    if (!threw) {
      std::printf("FAIL dod overflow\n");
      return 1;
    }
    e.Close();
    threw = false;
    try {
      e.Encode(m3tsz::Datapoint{0, 1}, S);
    } catch (const m3tsz::Error &err) {
      threw = err.status == M3TSZ_ERR_ENCODER_CLOSED;
    }
    if (!threw) return 1;
    std::printf("PASS\n");
    return 0;
  } catch (const m3tsz::Error &err) {
    if (err.status == M3TSZ_ERR_NO_DEVICE) {
      std::printf("NO_DEVICE\n");
      return 3;
    }
    std::printf("ERROR %s\n", err.what());
    return 2;
  }
}
