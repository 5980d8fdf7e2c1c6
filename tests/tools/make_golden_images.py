"""Golden fixtures for the host image codecs (sycl-ray-tracer_b200/host/image_codecs.hpp).

Inputs: small PNG / JPEG files written with Pillow in every variant the decoders claim (run here, where
Pillow exists). Outputs: what the REFERENCE's own vendored stb_image.h returns for
stbi_load_from_memory(bytes, size, &w, &h, &comp, 4) — compiled in place by `make -C oracle ref` into
oracle/_ref/libstbref.so — plus stbir_resize_uint8_srgb(... 512, 512 ..., STBIR_RGBA) of a few of them
(the reference's texture bake, src/image_manager.hpp:52-62). Stored in tests/golden/images.npz.

    python tests/tools/make_golden_images.py
"""
import ctypes as C
import io
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden", "images.npz")


RESIZE_PROCEDURAL = [(700, 600, 21), (1000, 100, 22), (515, 2048, 23), (512, 300, 24)]


def stbref():
    so = os.path.join(ROOT, "oracle", "_ref", "libstbref.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    L = C.CDLL(so)
    L.stbref_load_rgba.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int]
    L.stbref_resize_srgb_rgba.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.stbref_failure_reason.restype = C.c_char_p
    return L


def stb_decode(L, data):
    w, h, comp = C.c_int(), C.c_int(), C.c_int()
    out = np.zeros(1 << 22, np.uint8)
    if not L.stbref_load_rgba(data, len(data), C.byref(w), C.byref(h), C.byref(comp), out.ctypes.data, out.size):
        raise RuntimeError("stb_image failed: %s" % L.stbref_failure_reason().decode())
    return out[: w.value * h.value * 4].reshape(h.value, w.value, 4).copy(), comp.value


def stb_resize(L, rgba, n=512):
    h, w, _ = rgba.shape
    src = np.ascontiguousarray(rgba)
    out = np.zeros((n, n, 4), np.uint8)
    assert L.stbref_resize_srgb_rgba(src.ctypes.data, w, h, out.ctypes.data, n, n)
    return out


def test_image(w, h, seed, alpha=True):
    """smooth gradients + hard edges + noise: exercises DC/AC, chroma upsampling at edges, all filters"""
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 4), np.float64)
    img[..., 0] = 127 + 120 * np.sin(x / max(w, 1) * 5.0 + seed)
    img[..., 1] = 255 * y / max(h - 1, 1)
    img[..., 2] = 255 * ((x // 5 + y // 3) % 2)
    img[..., 3] = 255 * np.clip(1.2 - np.hypot(x - w / 2, y - h / 2) / max(w, h), 0, 1)
    img[..., :3] += rs.randint(-20, 21, (h, w, 3))
    if h > 6 and w > 6:
        img[2:5, 1:6, :3] = (255, 0, 0)
        img[-4:-1, -6:-2, :3] = (0, 0, 255)
    out = np.clip(img, 0, 255).astype(np.uint8)
    if not alpha:
        out[..., 3] = 255
    return out


def cases():
    from PIL import Image
    out = {}

    def save(name, im, fmt, **kw):
        b = io.BytesIO()
        im.save(b, fmt, **kw)
        out[name] = b.getvalue()

    rgba = test_image(37, 29, 1)
    rgb = Image.fromarray(rgba[..., :3], "RGB")
    # ---- PNG
    save("png_rgba8", Image.fromarray(rgba, "RGBA"), "PNG")
    save("png_rgb8", rgb, "PNG")
    save("png_grey8", rgb.convert("L"), "PNG")
    save("png_grey_alpha8", Image.fromarray(rgba, "RGBA").convert("LA"), "PNG")
    save("png_palette8", rgb.quantize(200), "PNG")
    save("png_palette4", rgb.quantize(16), "PNG", bits=4)
    save("png_palette2", rgb.quantize(4), "PNG", bits=2)
    save("png_palette1", rgb.quantize(2), "PNG", bits=1)
    pal = rgb.quantize(32)
    save("png_palette_trns", pal, "PNG", transparency=bytes([0, 80, 160, 255] * 8))
    save("png_grey1", rgb.convert("1"), "PNG")
    save("png_rgb8_colorkey", rgb, "PNG", transparency=(255, 0, 0))
    save("png_grey8_colorkey", rgb.convert("L"), "PNG", transparency=int(np.asarray(rgb.convert("L"))[3, 3]))
    g16 = (np.asarray(rgb.convert("L")).astype(np.uint16) * 257 + 13) & 0xFFFF
    save("png_grey16", Image.fromarray(g16.astype(np.uint16), "I;16"), "PNG")
    save("png_tiny_1x1", Image.fromarray(test_image(1, 1, 3), "RGBA"), "PNG")
    save("png_wide_3x1", Image.fromarray(test_image(3, 1, 4), "RGBA"), "PNG")
    # low-bit grey and interlaced variants are written by hand below (Pillow cannot emit them all)
    out.update(handmade_pngs())
    # ---- JPEG
    big = Image.fromarray(test_image(61, 45, 5)[..., :3], "RGB")
    for ss, tag in ((0, "444"), (1, "422"), (2, "420")):
        save(f"jpg_baseline_{tag}", big, "JPEG", quality=85, subsampling=ss)
        save(f"jpg_progressive_{tag}", big, "JPEG", quality=70, subsampling=ss, progressive=True)
    save("jpg_grey", big.convert("L"), "JPEG", quality=80)
    save("jpg_grey_progressive", big.convert("L"), "JPEG", quality=60, progressive=True)
    save("jpg_q100_420", big, "JPEG", quality=100, subsampling=2)
    save("jpg_q10_420", big, "JPEG", quality=10, subsampling=2)
    save("jpg_optimized", big, "JPEG", quality=75, optimize=True)
    save("jpg_restart", big, "JPEG", quality=75, subsampling=2, restart_marker_blocks=2)
    save("jpg_restart_progressive", big, "JPEG", quality=75, subsampling=1, restart_marker_blocks=3, progressive=True)
    save("jpg_tiny_1x1", Image.fromarray(test_image(1, 1, 6)[..., :3], "RGB"), "JPEG", quality=90)
    save("jpg_odd_9x17_420", Image.fromarray(test_image(9, 17, 7)[..., :3], "RGB"), "JPEG", quality=90, subsampling=2)
    save("jpg_odd_17x9_422", Image.fromarray(test_image(17, 9, 8)[..., :3], "RGB"), "JPEG", quality=90, subsampling=1)
    save("jpg_rgb_adobe", big, "JPEG", quality=90, subsampling=0, keep_rgb=True)
    save("jpg_noisy_128", Image.fromarray(np.random.RandomState(9).randint(0, 256, (128, 128, 3), dtype=np.uint8), "RGB"), "JPEG", quality=95, subsampling=2)
    # h1v2 and other sampling layouts through a raw re-tag of the SOF sampling factors are not valid
    # streams; the generic / v2 upsamplers are covered by jpg_odd cases plus the 4:4:0 file below
    try:
        save("jpg_440", big, "JPEG", quality=85, subsampling="4:4:0")
    except Exception:
        pass
    out.update(handmade_jpegs())
    return {k: v for k, v in out.items() if v is not None}


def handmade_pngs():
    """variants Pillow does not write: Adam7 interlace at every colour type / sub-byte depth, 2- and 4-bit grey,
    16-bit RGB(A), 16-bit colour key, all five filters on every row"""
    import struct
    import zlib

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    def pack_rows(samples, depth):
        """samples: (h, w*ch) integer array of channel values -> list of packed byte rows"""
        rows = []
        for r in samples:
            if depth == 8:
                rows.append(bytes(r.astype(np.uint8)))
            elif depth == 16:
                rows.append(r.astype(">u2").tobytes())
            else:
                bits = np.zeros(((len(r) * depth + 7) // 8) * 8, np.uint8)
                for i, v in enumerate(r):
                    for b in range(depth):
                        bits[i * depth + b] = (int(v) >> (depth - 1 - b)) & 1
                rows.append(np.packbits(bits).tobytes())
        return rows

    def filt(rows, bpp, mode):
        """apply PNG filters; mode 'cycle' uses filter y % 5 on row y"""
        out, prev = b"", None
        for y, row in enumerate(rows):
            ft = y % 5 if mode == "cycle" else 0
            cur, p = np.frombuffer(row, np.uint8).astype(np.int32), (np.frombuffer(prev, np.uint8).astype(np.int32) if prev is not None else np.zeros(len(row), np.int32))
            enc = np.zeros(len(row), np.int32)
            for x in range(len(row)):
                a = cur[x - bpp] if x >= bpp else 0
                b = p[x]
                c = p[x - bpp] if x >= bpp else 0
                if ft == 0:
                    pred = 0
                elif ft == 1:
                    pred = a
                elif ft == 2:
                    pred = b
                elif ft == 3:
                    pred = (a + b) // 2
                else:
                    pp = a + b - c
                    pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                enc[x] = (cur[x] - pred) & 255
            out += bytes([ft]) + bytes(enc.astype(np.uint8))
            prev = row
        return out

    def make(samples, w, h, ch, depth, ctype, interlace=False, extra=b"", mode="cycle"):
        s = samples.reshape(h, w, ch)
        bpp = max(1, ch * depth // 8)
        if not interlace:
            body = filt(pack_rows(s.reshape(h, w * ch), depth), bpp, mode)
        else:
            body = b""
            xo, yo, xs, ys = (0, 4, 0, 2, 0, 1, 0), (0, 0, 4, 0, 2, 0, 1), (8, 8, 4, 4, 2, 2, 1), (8, 8, 8, 4, 4, 2, 2)
            for p in range(7):
                sub = s[yo[p]::ys[p], xo[p]::xs[p]]
                if sub.shape[0] and sub.shape[1]:
                    body += filt(pack_rows(sub.reshape(sub.shape[0], -1), depth), bpp, mode)
        ihdr = struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + extra + chunk(b"IDAT", zlib.compress(body, 9)) + chunk(b"IEND", b"")

    rs = np.random.RandomState(11)
    out = {}
    w, h = 21, 19
    out["png_h_rgba8_adam7"] = make(rs.randint(0, 256, (h, w, 4)), w, h, 4, 8, 6, True)
    out["png_h_rgb8_adam7"] = make(rs.randint(0, 256, (h, w, 3)), w, h, 3, 8, 2, True)
    out["png_h_grey4"] = make(rs.randint(0, 16, (h, w, 1)), w, h, 1, 4, 0)
    out["png_h_grey2"] = make(rs.randint(0, 4, (h, w, 1)), w, h, 1, 2, 0)
    out["png_h_grey1_adam7"] = make(rs.randint(0, 2, (h, w, 1)), w, h, 1, 1, 0, True)
    out["png_h_grey4_adam7"] = make(rs.randint(0, 16, (h, w, 1)), w, h, 1, 4, 0, True)
    out["png_h_grey2_colorkey"] = make(rs.randint(0, 4, (h, w, 1)), w, h, 1, 2, 0, False, chunk(b"tRNS", struct.pack(">H", 2)))
    plte = chunk(b"PLTE", bytes(rs.randint(0, 256, 16 * 3).astype(np.uint8))) + chunk(b"tRNS", bytes(rs.randint(0, 256, 9).astype(np.uint8)))
    out["png_h_palette4_trns_adam7"] = make(rs.randint(0, 16, (h, w, 1)), w, h, 1, 4, 3, True, plte)
    out["png_h_rgb16"] = make(rs.randint(0, 65536, (h, w, 3)), w, h, 3, 16, 2)
    out["png_h_rgba16_adam7"] = make(rs.randint(0, 65536, (h, w, 4)), w, h, 4, 16, 6, True)
    s16 = rs.randint(0, 4, (h, w, 3)) * 21845
    out["png_h_rgb16_colorkey"] = make(s16, w, h, 3, 16, 2, False, chunk(b"tRNS", struct.pack(">HHH", 21845, 0, 43690)))
    out["png_h_grey_alpha16"] = make(rs.randint(0, 65536, (h, w, 2)), w, h, 2, 16, 4)
    out["png_h_small_adam7_2x3"] = make(rs.randint(0, 256, (3, 2, 4)), 2, 3, 4, 8, 6, True)
    return out


def handmade_jpegs():
    """sampling layouts Pillow cannot write (4:4:0, 4:1:1, mixed chroma factors, 3x / 4x factors): a minimal
    baseline encoder that entropy-codes RANDOM quantised coefficients (no forward DCT needed: any coefficient
    set is a valid image) with flat 8-bit Huffman tables, interleaved MCUs, optional restart markers."""
    import struct
    zig = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56,
           57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
    dc_syms = list(range(12))                                                    # 12 codes of 4 bits
    ac_syms = [0x00, 0xF0] + [(r << 4) | sz for r in range(16) for sz in range(1, 11)]   # 162 codes of 8 bits
    dc_code = {v: (i, 4) for i, v in enumerate(dc_syms)}
    ac_code = {v: (i, 8) for i, v in enumerate(ac_syms)}

    class Bits:
        def __init__(self):
            self.out, self.acc, self.n = bytearray(), 0, 0

        def put(self, code, length):
            self.acc = (self.acc << length) | (code & ((1 << length) - 1))
            self.n += length
            while self.n >= 8:
                b = (self.acc >> (self.n - 8)) & 255
                self.out.append(b)
                if b == 255:
                    self.out.append(0)
                self.n -= 8

        def flush(self):
            if self.n:
                self.put((1 << (8 - self.n)) - 1, 8 - self.n)

    def mag(v):
        sz = int(abs(v)).bit_length()
        return sz, (v if v >= 0 else v + (1 << sz) - 1)

    def make(w, h, factors, seed, restart=0, qscale=6):
        rs = np.random.RandomState(seed)
        hmax, vmax = max(f[0] for f in factors), max(f[1] for f in factors)
        mx, my = -(-w // (8 * hmax)), -(-h // (8 * vmax))
        seg = lambda m, d: b"\xff" + bytes([m]) + struct.pack(">H", len(d) + 2) + d
        out = b"\xff\xd8" + seg(0xE0, b"JFIF\0\1\1\0\0\1\0\1\0\0")
        q = bytes(int(min(255, qscale + (i // 8 + i % 8))) for i in range(64))
        out += seg(0xDB, b"\0" + q) + seg(0xDB, b"\1" + bytes(min(255, 2 * v) for v in q))
        out += seg(0xC0, struct.pack(">BHHB", 8, h, w, len(factors)) + b"".join(bytes([i + 1, (f[0] << 4) | f[1], 0 if i == 0 else 1]) for i, f in enumerate(factors)))
        out += seg(0xC4, b"\x00" + bytes([0, 0, 0, 12] + [0] * 12) + bytes(dc_syms))
        out += seg(0xC4, b"\x10" + bytes([0, 0, 0, 0, 0, 0, 0, 162] + [0] * 8) + bytes(ac_syms))
        if restart:
            out += seg(0xDD, struct.pack(">H", restart))
        out += seg(0xDA, bytes([len(factors)]) + b"".join(bytes([i + 1, 0x00]) for i in range(len(factors))) + b"\0\x3f\0")
        bits, pred, count, rst = Bits(), [0] * len(factors), 0, 0
        for _ in range(mx * my):
            for c, (fh, fv) in enumerate(factors):
                for _b in range(fh * fv):
                    coef = np.zeros(64, np.int64)
                    coef[0] = int(np.clip(pred[c] + rs.randint(-12, 13), -60, 60))
                    for k in rs.choice(np.arange(1, 64), rs.randint(0, 7), replace=False):
                        coef[k] = rs.randint(-9, 10) if k < 12 else rs.randint(-2, 3)
                    if rs.rand() < 0.1:
                        coef[rs.randint(40, 64)] = rs.randint(1, 3)             # long zero runs (ZRL)
                    sz, extra = mag(int(coef[0] - pred[c]))
                    pred[c] = int(coef[0])
                    bits.put(*dc_code[sz])
                    if sz:
                        bits.put(extra, sz)
                    run = 0
                    last = max([k for k in range(1, 64) if coef[zig[k]] != 0], default=0)
                    for k in range(1, last + 1):
                        v = int(coef[zig[k]])
                        if v == 0:
                            run += 1
                            continue
                        while run > 15:
                            bits.put(*ac_code[0xF0])
                            run -= 16
                        sz, extra = mag(v)
                        bits.put(*ac_code[(run << 4) | sz])
                        bits.put(extra, sz)
                        run = 0
                    if last < 63:
                        bits.put(*ac_code[0x00])
            count += 1
            if restart and count % restart == 0 and count < mx * my:
                bits.flush()
                bits.out += bytes([0xFF, 0xD0 + rst])
                bits.acc = bits.n = 0
                rst = (rst + 1) & 7
                pred = [0] * len(factors)
        bits.flush()
        return out + bytes(bits.out) + b"\xff\xd9"

    return {
        "jpg_h_440": make(45, 37, [(1, 2), (1, 1), (1, 1)], 1),                 # h1v2: vertical 3:1 filter
        "jpg_h_411": make(53, 21, [(4, 1), (1, 1), (1, 1)], 2),                 # 4x horizontal: replicated samples
        "jpg_h_mixed": make(40, 40, [(2, 2), (2, 1), (1, 2)], 3),               # Cb h1v2, Cr h2v1 against a 2x2 luma
        "jpg_h_4x4": make(37, 33, [(4, 4), (1, 1), (1, 1)], 4),                 # 32x32 MCUs
        "jpg_h_3x1": make(50, 9, [(3, 1), (1, 1), (1, 1)], 5),
        "jpg_h_42": make(70, 31, [(4, 2), (2, 1), (1, 1)], 6, restart=2),       # mixed ratios 1, 2 and 4 + restart markers
        "jpg_h_grey_odd": make(13, 29, [(1, 1)], 7),
        "jpg_h_chroma_major": make(33, 18, [(1, 1), (2, 2), (2, 2)], 8),        # luma subsampled instead of chroma
        "jpg_h_1px_wide": make(1, 40, [(2, 2), (1, 1), (1, 1)], 9),             # w_lores == 1 special cases
        "jpg_h_2px_422": make(2, 9, [(2, 1), (1, 1), (1, 1)], 10),
    }


def main():
    L = stbref()
    inputs = cases()
    store = {}
    for name, data in sorted(inputs.items()):
        rgba, comp = stb_decode(L, data)
        store["in_" + name] = np.frombuffer(data, np.uint8)
        store["out_" + name] = rgba
        store["comp_" + name] = np.int32(comp)
        print(f"{name:32s} {len(data):6d} bytes -> {rgba.shape[1]}x{rgba.shape[0]} comp {comp}")
    # the bake's resize (src/image_manager.hpp:52-62) of three decoded images, stored as every 8th texel of
    # every 8th row of the 512x512 result (64x64: enough to measure how far another filter is from it)
    for name in ("png_rgba8", "jpg_baseline_420", "jpg_noisy_128"):
        store["resize_" + name] = stb_resize(L, store["out_" + name])[::8, ::8].copy()
    # ... and of procedural inputs that are reduced (Mitchell kernel) or reduced along one axis only; the
    # input is test_image(w, h, seed), regenerated by the test
    for w, h, seed in RESIZE_PROCEDURAL:
        store[f"resize_proc_{w}x{h}_{seed}"] = stb_resize(L, test_image(w, h, seed))[::8, ::8].copy()
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
