"""Fileset files: the writer / reader of m3_b200/fileset.py against the
reference's format -- msgpack framing cross-checked with the independent `msgpack` package,
digest / checkpoint chain, per-entry checksums, corruption detection.  No GPU."""
import os
import struct
import zlib

import numpy as np
import pytest

from m3_b200 import fileset as F

SEC = 1_000_000_000


def _mk(tmp_path, n=50, seed=3):
    rng = np.random.default_rng(seed)
    ids = [b"series-%05d{host=h%d}" % (int(rng.integers(0, 10**5)), i) for i in range(n)]
    streams = [bytes(rng.integers(0, 256, size=int(rng.integers(1, 400)), dtype=np.uint8)) for _ in range(n)]
    # streams placed in arbitrary order with gaps, like the packed encoder's output
    order = rng.permutation(n)
    offsets, pos = [0] * n, 0
    for i in order:
        pos += int(rng.integers(0, 64))
        offsets[i] = pos
        pos += len(streams[i])
    blob = np.zeros(pos, dtype=np.uint8)
    for i in range(n):
        blob[offsets[i]: offsets[i] + len(streams[i])] = np.frombuffer(streams[i], dtype=np.uint8)
    tags = [b"tags%d" % i for i in range(n)]
    block_start = 1599955200 * SEC
    F.write_fileset(str(tmp_path), "metrics", 7, block_start, 7200 * SEC, ids, blob, offsets,
                    [len(s) for s in streams], tags, volume=2)
    return ids, streams, tags, block_start


def test_roundtrip_and_layout(tmp_path):
    ids, streams, tags, block_start = _mk(tmp_path)
    shard_dir = os.path.join(str(tmp_path), "data", "metrics", "7")
    names = sorted(os.listdir(shard_dir))
    assert names == sorted("fileset-%d-2-%s.db" % (block_start, s) for s in
                           ("info", "index", "summaries", "bloomfilter", "data", "digest", "checkpoint"))
    fs = F.read_fileset(str(tmp_path), "metrics", 7, block_start, volume=2)
    assert fs.info.block_start == block_start and fs.info.block_size == 7200 * SEC
    assert fs.info.entries == len(ids) and fs.info.volume_index == 2
    assert (fs.info.major_version, fs.info.minor_version) == (1, 1)
    assert fs.ids == sorted(ids)  # index entries are sorted by ID (write.go writeAll)
    by_id = dict(zip(ids, zip(streams, tags)))
    assert int(fs.sizes.sum()) == fs.data.shape[0]  # the data file is the segments back to back
    for k, id_ in enumerate(fs.ids):
        seg = bytes(fs.data[fs.offsets[k]: fs.offsets[k] + fs.sizes[k]])
        assert seg == by_id[id_][0] and fs.tags[k] == by_id[id_][1]
        assert fs.data_checksums[k] == zlib.adler32(seg)  # ts.Segment.CalculateChecksum


def test_msgpack_framing_matches_independent_decoder(tmp_path):
    msgpack = pytest.importorskip("msgpack")
    ids, streams, tags, block_start = _mk(tmp_path, n=9)
    p = lambda s: F.fileset_path(str(tmp_path), "metrics", 7, block_start, 2, s)
    objs = list(msgpack.Unpacker(open(p("info"), "rb"), raw=True))
    # version, [objectType, [11 fields]] (msgpack/encoder.go:280-293,380-384)
    assert objs[0] == 1 and objs[1][0] == 2 and len(objs[1][1]) == 11
    info = objs[1][1]
    assert info[0] == block_start and info[2] == 9 and info[3] == 1 and info[9] == 2 and info[10] == 1
    assert info[4] == [0] and info[5] == [0, 0]  # IndexSummariesInfo, IndexBloomFilterInfo
    objs = list(msgpack.Unpacker(open(p("index"), "rb"), raw=True))
    assert len(objs) == 2 * 9
    for k in range(9):
        ver, (typ, ent) = objs[2 * k: 2 * k + 2]
        assert ver == 1 and typ == 5 and len(ent) == 7 and ent[0] == k
        assert ent[2] == len(dict(zip(ids, streams))[ent[1]])
    # digest chain: five little-endian Adler-32s, checkpoint = digest of the digest file
    dig = open(p("digest"), "rb").read()
    assert len(dig) == 20
    for i, s in enumerate(("info", "index", "summaries", "bloomfilter", "data")):
        assert struct.unpack_from("<I", dig, 4 * i)[0] == zlib.adler32(open(p(s), "rb").read())
    assert struct.unpack("<I", open(p("checkpoint"), "rb").read())[0] == zlib.adler32(dig)


def test_corruption_is_detected(tmp_path):
    ids, streams, tags, block_start = _mk(tmp_path, n=12)
    p = lambda s: F.fileset_path(str(tmp_path), "metrics", 7, block_start, 2, s)
    raw = bytearray(open(p("index"), "rb").read())
    raw[len(raw) // 2] ^= 0x40
    open(p("index"), "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        F.read_fileset(str(tmp_path), "metrics", 7, block_start, volume=2)
    with pytest.raises(ValueError):  # the entry's own checksum (IndexEntry V3) also catches it
        F.read_fileset(str(tmp_path), "metrics", 7, block_start, volume=2, verify_digests=False)
    os.remove(p("checkpoint"))
    with pytest.raises(FileNotFoundError):
        F.read_fileset(str(tmp_path), "metrics", 7, block_start, volume=2)
