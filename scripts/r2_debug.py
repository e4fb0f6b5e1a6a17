import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
import oracle_lib as O
from m3_b200.codec import BatchCodec
gb = bytes([0x13, 0xce, 0x4c, 0xa4, 0x30, 0xcb, 0x40, 0x0,  0x80, 0x20, 0x1,  0x53, 0xe4,
            0x2,  0x80, 0x0,  0x0,  0x0,  0x0,  0x0,  0xb,  0xf1, 0x96, 0x6,  0x0,  0x81,
            0x0,  0x81, 0x68, 0x2,  0x1,  0x1,  0x0,  0x0,  0x0,  0x1d, 0xcd, 0x65, 0x0,
            0x0,  0x20, 0x8,  0x20, 0x18, 0x20, 0x2f, 0xf,  0xa6, 0x58, 0x77, 0x0,  0x80,
            0x40, 0x0,  0x0,  0x0,  0xe,  0xe6, 0xb2, 0x80, 0x23, 0x80, 0x0])
for int_opt in (False,):
    dps, err = O.decode_all(gb, int_opt)
    print("oracle", err, [(d[0], d[1], d[2]) for d in dps])
    codec = BatchCodec(0, int_opt)
    for cap in (2048, 8, 7):
        buf = torch.zeros(len(gb) + 16, dtype=torch.uint8, device="cuda")
        buf[:len(gb)] = torch.frombuffer(bytearray(gb), dtype=torch.uint8).cuda()
        r = codec.decode(buf[:len(gb)], torch.tensor([0, len(gb)], dtype=torch.int64, device="cuda"), cap)
        torch.cuda.synchronize()
        n = int(r.n_points[0])
        print("gpu cap", cap, int(r.status[0]), n, r.ts[0, :min(n, cap)].tolist(), r.values[0, :min(n, cap)].tolist())
    h_ts = torch.zeros((1, 2048), dtype=torch.int64); h_v = torch.zeros((1, 2048), dtype=torch.float64)
    h_n = torch.zeros(1, dtype=torch.int32); h_st = torch.zeros(1, dtype=torch.int32)
    codec.decode_host(torch.frombuffer(bytearray(gb), dtype=torch.uint8), torch.tensor([0, len(gb)], dtype=torch.int64),
                      2048, h_ts, h_v, h_n, h_st)
    n = int(h_n[0]); print("host", int(h_st[0]), n, h_ts[0, :n].tolist(), h_v[0, :n].tolist())
