/* m3tsz_segment_oracle.c -- TEST INFRASTRUCTURE (see m3tsz_oracle.h): CPU
 * restatement of the segment checksum the fileset read path verifies
 * (SURVEY.md section 8f row N2).
 *
 *   ts.Segment.CalculateChecksum      src/dbnode/ts/segment.go:60-76
 *       digest over Head then Tail == Adler-32 of the concatenated stream
 *   digest.Checksum                   src/dbnode/digest/digest.go:36-38  (hash/adler32)
 *   verification sites                src/dbnode/persist/fs/read.go:395-397 (streaming
 *       read), seek.go (errSeekChecksumMismatch), index entries carry DataChecksum
 *       (src/dbnode/persist/schema/types.go:70-78)
 *
 * Third-party: github.com/m3db/stackadler32 (go.mod:36) is an allocation-free
 * Adler-32 with the same result as Go's hash/adler32; the algorithm is the
 * published one (RFC 1950 section 8.2): a = 1 + sum(d_i) mod 65521,
 * b = sum of the running a's mod 65521, checksum = b << 16 | a.
 * Pinned in tests/test_checksum_oracle.py by the RFC / Go standard-library
 * known answers and by zlib.adler32 on random inputs. */
#include "m3tsz_oracle.h"

#define M3O_ADLER_MOD 65521u
/* largest n with 255 n (n+1) / 2 + (n+1)(65520) < 2^32 (zlib's NMAX) */
#define M3O_ADLER_NMAX 5552u

uint32_t m3o_adler32(const uint8_t *data, size_t n) {
  uint32_t a = 1, b = 0;
  while (n > 0) {
    size_t k = n < M3O_ADLER_NMAX ? n : M3O_ADLER_NMAX;
    n -= k;
    while (k--) {
      a += *data++;
      b += a;
    }
    a %= M3O_ADLER_MOD;
    b %= M3O_ADLER_MOD;
  }
  return (b << 16) | a;
}

/* checksum of every stream of a CSR batch; status[s] = 0, or M3O_ERR_CHECKSUM when
 * `expected` is given and differs (read.go:395-397) */
void m3o_adler32_batch(const uint8_t *streams, const uint64_t *offsets, uint64_t n_series,
                       const uint32_t *expected, uint32_t *out, int32_t *status) {
  for (uint64_t s = 0; s < n_series; s++) {
    const uint32_t c = m3o_adler32(streams + offsets[s], (size_t)(offsets[s + 1] - offsets[s]));
    if (out) out[s] = c;
    if (status) status[s] = (expected && expected[s] != c) ? M3O_ERR_CHECKSUM : 0;
  }
}
