"""read_data_files on the GPU: the reference's fileset inspection / benchmark tool
(/root/reference/src/cmd/tools/read_data_files/main/main.go:55-219) over the batch decoder.

    python -m m3_b200.tools.read_data_files -p /var/lib/m3db -n metrics -s 7 -b <blockStartNanos> \
        [-v volume] [-f id-substring] [-B series|datapoints] [--generate N_SERIES]

Same flags and the same report as the reference tool.  The whole volume is read with ONE
upload of the data file (it already is the decoder's input layout), ONE Adler-32 launch that
checks every segment against its index entry (persist/fs/read.go:395-397 does it per entry)
and ONE decode launch.  `--generate N` first writes a synthetic volume (Gaussian-walk series,
1440 points, encoded on the GPU, written in the reference's fileset format) so the benchmark
can run where no M3DB data directory exists.
"""
import argparse
import base64
import os
import sys
import time

import numpy as np


def _generate(args, codec):
    import torch
    from .. import fileset, synth
    S, P = args.generate, args.points
    ts, vals, start = synth.gaussian_walk(S, P, codec.device, seed=args.seed)
    pk = codec.encode_packed(ts, vals, start, unit=1, align=1)
    torch.cuda.synchronize()
    total = int(pk.total.item())
    ids = [b"synthetic.series.%07d" % i for i in range(S)]
    block_start = int(start[0].item()) if args.block_start <= 0 else args.block_start
    fileset.write_fileset(args.path_prefix, args.namespace, max(args.shard, 0), block_start, P * 60 * 10**9, ids,
                          pk.packed[:total].cpu().numpy(), pk.offsets.cpu().numpy(), pk.out_len.cpu().numpy(),
                          volume=args.volume)
    return block_start


def main(argv=None):
    ap = argparse.ArgumentParser(prog="read_data_files", description=__doc__.split("\n\n")[0])
    ap.add_argument("-p", "--path-prefix", required=True, help="Path prefix [e.g. /var/lib/m3db]")
    ap.add_argument("-n", "--namespace", default="default", help="Namespace [e.g. metrics]")
    ap.add_argument("-s", "--shard", type=int, default=-1, help="Shard, or -1 for all shards in the directory")
    ap.add_argument("-b", "--block-start", type=int, default=0, help="Block Start Time [in nsec]")
    ap.add_argument("-v", "--volume", type=int, default=0, help="Volume number")
    ap.add_argument("-t", "--fileset-type", default="flush", choices=["flush"], help="flush (snapshots: not built)")
    ap.add_argument("-f", "--id-filter", default="", help="ID Contains Filter (optional)")
    ap.add_argument("-B", "--benchmark", default="", choices=["", "series", "datapoints"],
                    help="benchmark mode (optional), [series|datapoints]")
    ap.add_argument("--generate", type=int, default=0, help="write a synthetic volume of N series first")
    ap.add_argument("--points", type=int, default=1440)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-points", type=int, default=0, help="decode capacity per series (0 = block size / 1s, capped)")
    ap.add_argument("--no-int-optimized", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)

    import torch
    from .. import fileset
    from ..codec import BatchCodec
    codec = BatchCodec(args.device, int_optimized=not args.no_int_optimized)
    if args.generate:
        args.block_start = _generate(args, codec)
        print("generated %d series x %d points at block start %d" % (args.generate, args.points, args.block_start))
    if args.block_start <= 0:
        ap.error("--block-start is required")
    if args.shard < 0:
        ns_dir = os.path.join(args.path_prefix, "data", args.namespace)
        shards = sorted(int(d) for d in os.listdir(ns_dir) if d.isdigit())
    else:
        shards = [args.shard]

    for shard in shards:
        start = time.perf_counter()
        fs = fileset.read_fileset(args.path_prefix, args.namespace, shard, args.block_start, args.volume)
        n = len(fs.ids)
        keep = np.arange(n)
        if args.id_filter:
            f = args.id_filter.encode()
            keep = np.array([i for i, id_ in enumerate(fs.ids) if f in id_], dtype=np.int64)
        series_count, datapoint_count, annotation_total = len(keep), 0, 0
        if args.benchmark != "series" and len(keep):
            cap = args.max_points or max(16, min(1 << 16, args.points if args.generate else 4096))
            sub = fs if len(keep) == n else fileset.FilesetData(
                fs.info, [fs.ids[i] for i in keep], [fs.tags[i] for i in keep], fs.offsets[keep], fs.sizes[keep],
                fs.data_checksums[keep], fs.data)
            while True:
                res, ck = fileset.decode_fileset(codec, sub, cap, want_events=(0 if args.benchmark else 1 << 16))
                torch.cuda.synchronize()
                st = res.status.cpu().numpy()
                npts = res.n_points.cpu().numpy().astype(np.int64)
                if (st == 100).any():  # M3TSZ_ERR_CAPACITY: a longer block than guessed
                    cap = int(npts.max())
                    continue
                break
            bad_ck = np.nonzero(ck.cpu().numpy())[0] if ck is not None else []
            if len(bad_ck):
                sys.exit("checksum does not match expected checksum for %r" % sub.ids[int(bad_ck[0])])
            bad = np.nonzero(st)[0]
            if len(bad):
                sys.exit("unable to iterate original data: status %d for %r" % (st[bad[0]], sub.ids[int(bad[0])]))
            datapoint_count = int(npts.sum())
            if not args.benchmark:  # print every datapoint like the reference tool
                tsv, vv = res.ts.cpu().numpy(), res.values.cpu().numpy()
                anns = {}
                if res.event_count is not None:
                    from .. import capi
                    ne = min(int(res.event_count.item()), res.events.shape[0])
                    raw = res.events[:ne].cpu().numpy().tobytes()
                    for e in (capi.DpEvent * ne).from_buffer_copy(raw):
                        if e.kind == capi.EVENT_ANNOTATION:
                            anns[(int(e.series), int(e.dp_index))] = (int(e.bit_offset), int(e.length))
                for k in range(len(sub.ids)):
                    seg = bytes(sub.data[sub.offsets[k]: sub.offsets[k] + sub.sizes[k]])
                    for i in range(int(npts[k])):
                        line = "{id: %s, dp: {TimestampNanos:%d Value:%s}" % (sub.ids[k].decode("utf-8", "replace"),
                                                                               tsv[k, i], repr(float(vv[k, i])))
                        if (k, i) in anns:
                            bo, ln = anns[(k, i)]
                            bits = int.from_bytes(seg, "big")
                            a = bytes((bits >> (len(seg) * 8 - bo - 8 * (j + 1))) & 0xFF for j in range(ln))
                            annotation_total += ln
                            line += ", annotation: %s" % base64.b64encode(a).decode()
                        print(line + "}")
        if series_count != fs.info.entries and not args.id_filter:
            print("actual time series count (%d) did not match info file data (%d)" % (series_count, fs.info.entries),
                  file=sys.stderr)
        if args.benchmark:
            run = time.perf_counter() - start
            print("Running time: %.6fs" % run)
            print("\n%d series read" % series_count)
            if run > 0:
                print("(%.2f series/second)" % (series_count / run))
            if args.benchmark == "datapoints":
                print("\n%d datapoints decoded" % datapoint_count)
                if run > 0:
                    print("(%.2f datapoints/second)" % (datapoint_count / run))
                print("\nTotal annotation size: %d bytes" % annotation_total)
                print("(data file %d bytes: %.2f GB/s from disk bytes, %d launches)"
                      % (fs.data.shape[0], fs.data.shape[0] / run / 1e9, codec.launch_count()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
