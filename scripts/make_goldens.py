#!/usr/bin/env python3
"""Transcribes the reference's own golden vectors for the M3TSZ path into
tests/golden/m3tsz_goldens.json.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):   python scripts/make_goldens.py

Every vector records the reference test file:line it was taken from.  Byte
arrays and datapoint tables are transcribed by hand below (they are test DATA,
not implementation); the long base64 fixtures are pulled out of the Go test
files mechanically so they cannot be mistyped.
"""
import json
import os
import re

REF = "/root/reference/src/dbnode/encoding"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "m3tsz_goldens.json")

SEC = 1_000_000_000
MS = 1_000_000
US = 1_000
S, MSU, USU, NSU = 1, 2, 3, 4  # xtime.Unit values (src/x/time/unit.go:30-42)

TEST_START = 1427162400 * SEC      # encoder_test.go:42
DP_START = 1427162462 * SEC        # encoder_test.go:216


def hx(bs):
    return bytes(bs).hex()


def go_strings(path, var):
    src = open(path).read()
    m = re.search(r"var %s = \[\]string\{(.*?)\n\}" % var, src, re.S)
    return re.findall(r'"([A-Za-z0-9+/=]+)"', m.group(1))


def main():
    g = {}

    # ---- bit I/O: ostream_test.go:44-71 ----
    g["ostream_write_bits"] = {
        "src": "encoding/ostream_test.go:44-71",
        "steps": [
            {"value": v, "nbits": n, "bytes": hx(b), "pos": p}
            for v, n, b, p in [
                (0x1, 1, [0x80], 1),
                (0x4, 3, [0xc0], 4),
                (0xa, 4, [0xca], 8),
                (0xfe, 8, [0xca, 0xfe], 8),
                (0xaafe, 7, [0xca, 0xfe, 0xfc], 7),
                (0x3, 2, [0xca, 0xfe, 0xfd, 0x80], 1),
                (0x1234567890abcdef, 64,
                 [0xca, 0xfe, 0xfd, 0x89, 0x1a, 0x2b, 0x3c, 0x48, 0x55, 0xe6, 0xf7, 0x80], 1),
                (0x1, 0,
                 [0xca, 0xfe, 0xfd, 0x89, 0x1a, 0x2b, 0x3c, 0x48, 0x55, 0xe6, 0xf7, 0x80], 1),
                (0x1, 65,
                 [0xca, 0xfe, 0xfd, 0x89, 0x1a, 0x2b, 0x3c, 0x48, 0x55, 0xe6, 0xf7, 0x80, 0x0, 0x0,
                  0x0, 0x0, 0x0, 0x0, 0x0, 0x80], 1),
            ]
        ],
    }
    # ---- istream_test.go:32-186 ----
    g["istream_read_bits"] = {
        "src": "encoding/istream_test.go:32-52",
        "bytes": hx([0xca, 0xfe, 0xfd, 0x89, 0x1a, 0x2b, 0x3c, 0x48, 0x55, 0xe6, 0xf7, 0x80, 0x0,
                     0x0, 0x0, 0x0, 0x0, 0x0, 0x0, 0x80]),
        "nbits": [1, 3, 4, 8, 7, 2, 64, 64],
        "expected": [0x1, 0x4, 0xa, 0xfe, 0x7e, 0x3, 0x1234567890abcdef, 0x1],
        "then_read_8_is_eof": True,
    }
    g["istream_peek_bits"] = {
        "src": "encoding/istream_test.go:76-100",
        "bytes": hx([0xa9, 0xfe, 0xfe, 0xdf, 0x9b, 0x57, 0x21, 0xf1]),
        "cases": [[0, 0], [1, 0x1], [8, 0xa9], [10, 0x2a7], [13, 0x153f], [16, 0xa9fe],
                  [32, 0xa9fefedf], [64, 0xa9fefedf9b5721f1]],
    }
    g["istream_peek_error"] = {"src": "encoding/istream_test.go:102-108", "bytes": "0102",
                               "peek": 20}
    g["istream_read_after_peek"] = {
        "src": "encoding/istream_test.go:110-133", "bytes": "abcd",
        "peek10": 0x2af, "peek20_eof": True, "reads": [[2, 0x2], [9, 0x15e]], "then_read_8_is_eof": True,
    }
    g["istream_peek_after_read"] = {
        "src": "encoding/istream_test.go:135-162",
        "bytes": hx([0x1, 0x2, 0x3, 0x4, 0x5, 0x6, 0x7, 0x8, 0x9, 0xA]),
        "ops": [["read", 16, 0x102], ["peek", 63, 0x30405060708090A >> 1],
                ["peek", 64, 0x30405060708090A], ["read", 1, 0], ["peek", 63, 0x30405060708090A],
                ["peek_eof", 64, 0]],
    }

    # ---- field level: encoder_test.go:54-123 ----
    g["write_dod_unit_unchanged"] = {
        "src": "encoding/m3tsz/encoder_test.go:54-81",
        "cases": [
            {"delta_ns": d, "unit": u, "bytes": hx(b), "pos": p}
            for d, u, b, p in [
                (0, S, [0x0], 1),
                (32 * SEC, S, [0x90, 0x0], 1),
                (-63 * SEC, S, [0xa0, 0x80], 1),
                (-128 * SEC, S, [0xd8, 0x0], 4),
                (255 * SEC, S, [0xcf, 0xf0], 4),
                (-2048 * SEC, S, [0xe8, 0x0], 8),
                (2047 * SEC, S, [0xe7, 0xff], 8),
                (4096 * SEC, S, [0xf0, 0x0, 0x1, 0x0, 0x0], 4),
                (-4096 * SEC, S, [0xff, 0xff, 0xff, 0x0, 0x0], 4),
                (4096 * SEC, NSU, [0xf0, 0x0, 0x0, 0x3b, 0x9a, 0xca, 0x0, 0x0, 0x0], 4),
                (-4096 * SEC, NSU, [0xff, 0xff, 0xff, 0xc4, 0x65, 0x36, 0x0, 0x0, 0x0], 4),
            ]
        ],
    }
    g["write_dod_unit_changed"] = {
        "src": "encoding/m3tsz/encoder_test.go:83-101",
        "cases": [
            {"delta_ns": d, "bytes": hx(b), "pos": p}
            for d, b, p in [
                (0, [0] * 8, 8),
                (32 * MS, [0x0, 0x0, 0x0, 0x0, 0x1, 0xe8, 0x48, 0x0], 8),
                (-63 * US, [0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0x9, 0xe8], 8),
            ]
        ],
    }
    g["write_xor"] = {
        "src": "encoding/m3tsz/encoder_test.go:103-123",
        "cases": [
            {"prev_xor": a, "cur_xor": b, "bytes": hx(c), "pos": p}
            for a, b, c, p in [
                (0x4028000000000000, 0, [0x0], 1),
                (0x4028000000000000, 0x0120000000000000, [0x80, 0x90], 6),
                (0x0120000000000000, 0x4028000000000000, [0xc1, 0x2e, 0x1, 0x40], 2),
            ]
        ],
    }
    # TestWriteAnnotation encoder_test.go:125-155 (NewTimestampEncoder(0, ns); writeAnnotation)
    g["write_annotation"] = {
        "src": "encoding/m3tsz/encoder_test.go:125-155",
        "cases": [
            {"annotation": "", "bytes": "", "pos": 0},
            {"annotation": "0102", "bytes": hx([0x80, 0x20, 0x40, 0x20, 0x40]), "pos": 3},
            {"annotation": "ff" * 8,
             "bytes": hx([0x80, 0x21, 0xdf, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xe0]), "pos": 3},
        ],
    }
    # TestWriteTimeUnit encoder_test.go:172-205 (TimeUnit forced to None before the call)
    g["write_time_unit"] = {
        "src": "encoding/m3tsz/encoder_test.go:172-205",
        "cases": [
            {"unit": 0, "result": False, "bytes": "", "pos": 0},
            {"unit": S, "result": True, "bytes": "804020", "pos": 3},
            {"unit": 255, "result": False, "bytes": "", "pos": 0},
        ],
    }
    # TestInitTimeUnit encoder_test.go:395-410
    g["init_time_unit"] = {
        "src": "encoding/m3tsz/encoder_test.go:395-410",
        "cases": [
            {"start_ns": 1, "unit": S, "expected": 0},
            {"start_ns": 1 * SEC, "unit": S, "expected": S},
            {"start_ns": 1 * SEC, "unit": 0, "expected": 0},
        ],
    }

    # ---- full streams (intOptimized=false): encoder_test.go:207-393 <-> iterator_test.go:181-385 ----
    def dps(rows):
        return [{"ts": DP_START + off, "value": v, "unit": u, "annotation": a} for off, v, u, a in rows]

    g["streams"] = [
        {
            "name": "no_annotation",
            "src": "encoding/m3tsz/encoder_test.go:207-245, iterator_test.go:181-222",
            "int_optimized": False, "encoder_start": TEST_START,
            "datapoints": dps([
                (0, 12, S, ""), (60 * SEC, 12, S, ""), (120 * SEC, 24, S, ""),
                (-76 * SEC, 24, S, ""), (-16 * SEC, 24, S, ""), (2092 * SEC, 15, S, ""),
                (4200 * SEC, 12, S, "")]),
            "decoded_annotations": ["", "", "", "", "", "", ""],
            "bytes": hx([
                0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0, 0x9f, 0x20, 0x14, 0x0, 0x0,
                0x0, 0x0, 0x0, 0x0, 0x5f, 0x8c, 0xb0, 0x3a, 0x0, 0xe1, 0x0, 0x78, 0x0, 0x0,
                0x40, 0x6, 0x58, 0x76, 0x8e, 0x0, 0x0]),
            "raw": hx([
                0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0, 0x9f, 0x20, 0x14, 0x0, 0x0,
                0x0, 0x0, 0x0, 0x0, 0x5f, 0x8c, 0xb0, 0x3a, 0x0, 0xe1, 0x0, 0x78, 0x0, 0x0,
                0x40, 0x6, 0x58, 0x76, 0x8c]),
            "raw_pos": 6,
        },
        {
            "name": "with_annotation",
            "src": "encoding/m3tsz/encoder_test.go:247-289, iterator_test.go:224-269",
            "int_optimized": False, "encoder_start": TEST_START,
            "datapoints": dps([
                (0, 12, S, "0a"), (60 * SEC, 12, S, "0a"), (120 * SEC, 24, S, ""),
                (-76 * SEC, 24, S, ""), (-16 * SEC, 24, S, "0102"), (2092 * SEC, 15, S, ""),
                (4200 * SEC, 12, S, "")]),
            "decoded_annotations": ["0a", "", "", "", "0102", "", ""],
            "bytes": hx([
                0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0, 0x80, 0x20, 0x1, 0x53, 0xe4,
                0x2, 0x80, 0x0, 0x0, 0x0, 0x0, 0x0, 0xb, 0xf1, 0x96, 0x7, 0x40, 0x10, 0x4,
                0x8, 0x4, 0xb, 0x84, 0x1, 0xe0, 0x0, 0x1, 0x0, 0x19, 0x61, 0xda, 0x38, 0x0]),
            "raw": hx([
                0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0, 0x80, 0x20, 0x1, 0x53, 0xe4,
                0x2, 0x80, 0x0, 0x0, 0x0, 0x0, 0x0, 0xb, 0xf1, 0x96, 0x7, 0x40, 0x10, 0x4,
                0x8, 0x4, 0xb, 0x84, 0x1, 0xe0, 0x0, 0x1, 0x0, 0x19, 0x61, 0xda, 0x30]),
            "raw_pos": 4,
        },
        {
            "name": "with_time_unit",
            "src": "encoding/m3tsz/encoder_test.go:291-327, iterator_test.go:271-313",
            "int_optimized": False, "encoder_start": TEST_START,
            "datapoints": dps([
                (0, 12, S, ""), (60 * SEC, 12, S, ""), (120 * SEC, 24, S, ""),
                (-76 * SEC, 24, S, ""), (-16 * SEC, 24, S, ""),
                (-15500000000, 15, NSU, ""), (-1400 * MS, 12, MSU, ""),
                (-10 * SEC, 12, S, ""), (10 * SEC, 12, S, "")]),
            "decoded_annotations": [""] * 9,
            "bytes": hx([
                0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0, 0x9f, 0x20, 0x14, 0x0, 0x0,
                0x0, 0x0, 0x0, 0x0, 0x5f, 0x8c, 0xb0, 0x3a, 0x0, 0xe1, 0x0, 0x40, 0x20,
                0x4f, 0xff, 0xff, 0xff, 0x22, 0x58, 0x60, 0xd0, 0xc, 0xb0, 0xee, 0x1, 0x1,
                0x0, 0x0, 0x0, 0x1, 0xa4, 0x36, 0x76, 0x80, 0x47, 0x0, 0x80, 0x7f, 0xff,
                0xff, 0xff, 0x7f, 0xd9, 0x9a, 0x80, 0x11, 0x44, 0x0]),
        },
        {
            "name": "with_annotation_and_time_unit",
            "src": "encoding/m3tsz/encoder_test.go:329-393, iterator_test.go:315-385",
            "int_optimized": False, "encoder_start": TEST_START,
            "datapoints": dps([
                (0, 12, S, "0a"), (60 * SEC, 12, S, ""), (120 * SEC, 24, S, ""),
                (-76 * SEC, 24, S, "0102"), (-16 * SEC, 24, MSU, ""),
                (-15500 * MS, 15, MSU, "030405"), (-14000 * MS, 12, S, "")]),
            "decoded_annotations": ["0a", "", "", "0102", "", "030405", ""],
            "bytes": hx([
                0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0, 0x80, 0x20, 0x1, 0x53, 0xe4,
                0x2, 0x80, 0x0, 0x0, 0x0, 0x0, 0x0, 0xb, 0xf1, 0x96, 0x6, 0x0, 0x81, 0x0,
                0x81, 0x68, 0x2, 0x1, 0x1, 0x0, 0x0, 0x0, 0x1d, 0xcd, 0x65, 0x0, 0x0, 0x20,
                0x8, 0x20, 0x18, 0x20, 0x2f, 0xf, 0xa6, 0x58, 0x77, 0x0, 0x80, 0x40, 0x0,
                0x0, 0x0, 0xe, 0xe6, 0xb2, 0x80, 0x23, 0x80, 0x0]),
        },
    ]

    # ---- iterator field level: iterator_test.go:44-179 ----
    g["read_next_timestamp"] = {
        "src": "encoding/m3tsz/iterator_test.go:44-88",
        "cases": [
            {"prev_delta_ns": pd, "unit": u, "bytes": hx(b), "expected_delta_ns": ed}
            for pd, u, b, ed in [
                (62 * SEC, S, [0x0], 62 * SEC),
                (65 * SEC, S, [0xa0, 0x0], 1 * SEC),
                (65 * SEC, S, [0x90, 0x0], 97 * SEC),
                (65 * SEC, S, [0xd0, 0x0], -191 * SEC),
                (65 * SEC, S, [0xcf, 0xf0], 320 * SEC),
                (65 * SEC, S, [0xe8, 0x0], -1983 * SEC),
                (65 * SEC, S, [0xe7, 0xff], 2112 * SEC),
                (65 * SEC, S, [0xf0, 0x0, 0x1, 0x0, 0x0], 4161 * SEC),
                (65 * SEC, S, [0xff, 0xff, 0xff, 0x0, 0x0], -4031 * SEC),
                (65 * SEC, NSU, [0xff, 0xff, 0xff, 0xc4, 0x65, 0x36, 0x0, 0x0, 0x0], -4031 * SEC),
                (65 * SEC, S, [0x80, 0x40, 0x40, 0x0, 0x0, 0x0, 0x0, 0x0, 0x0, 0x7d, 0x0],
                 65000001 * US),
            ]
        ],
        "error_stream": "01",  # readFirstTimestamp, then readNextTimestamp x2 all error
    }
    g["read_next_value"] = {
        "src": "encoding/m3tsz/iterator_test.go:90-115",
        "cases": [
            {"prev_value": pv, "prev_xor": px, "bytes": hx(b), "expected_xor": ex, "expected_value": ev}
            for pv, px, b, ex, ev in [
                (0x1234, 0x4028000000000000, [0x0], 0x0, 0x1234),
                (0xaaaaaa, 0x4028000000000000, [0x80, 0x90], 0x0120000000000000, 0x0120000000aaaaaa),
                (0xdeadbeef, 0x0120000000000000, [0xc1, 0x2e, 0x1, 0x40], 0x4028000000000000,
                 0x40280000deadbeef),
            ]
        ],
        "error_stream": "f0",
    }
    g["read_annotation"] = {
        "src": "encoding/m3tsz/iterator_test.go:117-149",
        "cases": [
            {"bytes": "00ff", "annotation": "ff"},
            {"bytes": "020203", "annotation": "0203"},
            {"bytes": "0e" + "ff" * 8, "annotation": "ff" * 8},
            {"bytes": "10" + "ff" * 9, "annotation": "ff" * 9},
        ],
    }
    g["read_time_unit"] = {
        "src": "encoding/m3tsz/iterator_test.go:151-179",
        "cases": [
            {"unit": MSU, "bytes": "01", "expected_unit": S, "expected_changed": True},
            {"unit": S, "bytes": "00", "expected_unit": 0, "expected_changed": False},
        ],
    }
    g["iterator_error_streams"] = {
        "src": "encoding/m3tsz/iterator_test.go:265-269,387-394",
        "cases": [
            {"bytes": hx([0x0, 0x0, 0x0, 0x0, 0x55, 0x10, 0xc5, 0x20, 0x80, 0x20, 0x1, 0x50, 0x8]),
             "int_optimized": False, "note": "annotation truncated: Next false, not done, has error"},
            {"bytes": hx([0x0, 0x0, 0x0, 0x0, 0x55, 0x10, 0xc5, 0x20, 0x80, 0x41, 0x20, 0x0, 0x0]),
             "int_optimized": False, "note": "unexpected time unit: Next false, Err != nil"},
        ],
    }

    # ---- DoD overflow error cases: encoder_test.go:581-730 (values checked below) ----
    g["fixtures_b64"] = {
        "src": "encoding/m3tsz/encoder_benchmark_test.go:36-47 (intOptimized=true)",
        "streams": go_strings(os.path.join(REF, "m3tsz", "encoder_benchmark_test.go"),
                              "sampleSeriesBase64"),
        "expected_points": [720, 720, 719, 720, 719, 720, 719, 720, 720, 720],
    }
    src = open(os.path.join(REF, "m3tsz", "iterator_test.go")).read()
    m = re.search(r'b64 := "([A-Za-z0-9+/=]+)"', src)
    g["regression_b64"] = {
        "src": "encoding/m3tsz/iterator_test.go:396-412 (intOptimized=true; first value -2^63, sig=64)",
        "stream": m.group(1),
    }

    # ---- roundtrip_test.go:270-285 overflow datapoints (property: round-trips in both modes) ----
    large = float(2 ** 63)  # float64(math.MaxInt64 - 1)
    g["overflow_values"] = {
        "src": "encoding/m3tsz/roundtrip_test.go:270-285",
        "values": [large, 10, -large, 10, -large, large, -12, large, 14.5, large, -large,
                   12.34858499392, large],
        "start_ns": DP_START, "step_ns": SEC,
    }
    g["precision_value"] = {"src": "encoding/m3tsz/roundtrip_test.go:94-106", "value": 187.80131100000006}

    # ---- aggregation.Gauge tables: src/aggregator/aggregation/gauge_test.go ----
    # time.Now() in the Go tests is strictly increasing between calls: times are given as offsets.
    g["gauge"] = {
        "src": "aggregator/aggregation/gauge_test.go:35-57 (default), :59-105 (UpdatePrevious prefix), "
               ":117-170 (custom types), :173-195 (last, out of order)",
        "cases": [
            {"name": "1..100", "src_line": "gauge_test.go:35-48,121-147",
             "times": list(range(100)), "values": [float(i) for i in range(1, 101)],
             "last": 100.0, "count": 100, "mean": 50.5, "max": 100.0, "min": 1.0, "sum": 5050.0},
            {"name": "empty", "src_line": "gauge_test.go:49-57,149-170",
             "times": [], "values": [],
             "last": 0.0, "count": 0, "mean": 0.0, "max": "NaN", "min": "NaN", "sum": 0.0},
            {"name": "1,2,3", "src_line": "gauge_test.go:64-75",
             "times": [0, 1, 2], "values": [1.0, 2.0, 3.0],
             "last": 3.0, "count": 3, "mean": 2.0, "max": 3.0, "min": 1.0, "sum": 6.0},
            {"name": "last out of order", "src_line": "gauge_test.go:173-195",
             # timeMid = now+60s; pre = mid-1s; prepre = mid-1s; after = mid+1s
             "times": [60, 59, 61, 59], "values": [42.0, 41.0, 43.0, 40.0],
             "last": 43.0, "out_of_order": 2},  # the test asserts only Last() and the counter
        ],
    }

    # ---- Prometheus conversion epilogue: src/query/storage/prom_converter_test.go ----
    HOUR, MIN = 3600 * SEC, 60 * SEC
    T0 = 1600000000 // 3600 * 3600 * SEC  # xtime.Now().Truncate(time.Hour): any hour-aligned instant

    def dps(*pairs):
        return [[T0 + off, v] for off, v in pairs]

    def samples(*pairs):
        return [[(T0 + off) // MS, v] for off, v in pairs]

    g["prom_counter_normalization"] = {
        "src": "query/storage/prom_converter_test.go:319-440 (resolutionThreshold = 5m default, options.go:30; "
               "handleResets = isCounter && maxResolution >= threshold, prom_converter.go:74-82)",
        "resolution_threshold_ns": 5 * MIN,
        "cases": [
            {"name": "low resolution gauge", "is_counter": False, "max_resolution_ns": HOUR,
             "given": dps((0, 1), (HOUR, 2)), "want": samples((0, 1), (HOUR, 2))},
            {"name": "high resolution gauge", "is_counter": False, "max_resolution_ns": MIN,
             "given": dps((0, 1), (MIN, 2)), "want": samples((0, 1), (MIN, 2))},
            {"name": "low resolution counter, no datapoints", "is_counter": True, "max_resolution_ns": HOUR,
             "given": [], "want": []},
            {"name": "low resolution counter, one datapoint", "is_counter": True, "max_resolution_ns": HOUR,
             "given": dps((0, 1)), "want": samples((0, 1))},
            {"name": "high resolution counter with no resets", "is_counter": True, "max_resolution_ns": MIN,
             "given": dps((0, 1), (MIN, 2), (2 * MIN, 2), (3 * MIN, 3)),
             "want": samples((0, 1), (MIN, 2), (2 * MIN, 2), (3 * MIN, 3))},
            {"name": "high resolution counter with resets", "is_counter": True, "max_resolution_ns": MIN,
             "given": dps((0, 10), (MIN, 3), (2 * MIN, 5), (3 * MIN, 8)),
             "want": samples((0, 10), (MIN, 3), (2 * MIN, 5), (3 * MIN, 8))},
            {"name": "low resolution counter with no resets", "is_counter": True, "max_resolution_ns": HOUR,
             "given": dps((0, 1), (MIN, 2), (HOUR, 2), (HOUR + MIN, 3)),
             "want": samples((MIN, 2), (HOUR + MIN, 3))},
            {"name": "low resolution counter with resets", "is_counter": True, "max_resolution_ns": HOUR,
             "given": dps((0, 10), (MIN, 3), (HOUR, 5), (HOUR + MIN, 8)),
             "want": samples((MIN, 13), (HOUR + MIN, 18))},
        ],
    }
    a, b = 187.80131100000006, 187.801311
    g["prom_value_decrease_tolerance"] = {
        "src": "query/storage/prom_converter_test.go:444-500 (datapoint i at now + i minutes)",
        "now_ns": T0, "step_ns": MIN,
        "cases": [
            {"name": "no tolerance", "given": [a, b, a, b, 200, 199.99], "tolerance": 0, "until_ns": 0,
             "want": [a, b, a, b, 200, 199.99]},
            {"name": "low tolerance", "given": [a, b, a, b, 200, 199.99], "tolerance": 0.00000001,
             "until_ns": T0 + HOUR, "want": [a, a, a, a, 200, 199.99]},
            {"name": "high tolerance", "given": [a, b, a, b, 200, 199.99], "tolerance": 0.0001,
             "until_ns": T0 + HOUR, "want": [a, a, a, a, 200, 200]},
            {"name": "tolerance expired", "given": [200, 199.99, 200, 199.99, 200, 199.99], "tolerance": 0.0001,
             "until_ns": T0, "want": [200, 199.99, 200, 199.99, 200, 199.99]},
            {"name": "tolerance expires in the middle", "given": [200, 199.99, 200, 199.99, 200, 199.99],
             "tolerance": 0.0001, "until_ns": T0 + 3 * MIN, "want": [200, 200, 200, 199.99, 200, 199.99]},
        ],
    }

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT, "fixtures:", len(g["fixtures_b64"]["streams"]))


if __name__ == "__main__":
    main()
