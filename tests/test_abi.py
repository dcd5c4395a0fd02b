"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every
symbol the headers declare, and refuses to work without a GPU (no CPU fallback)."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: str):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_(?:cuda|scene|make)_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    """One library per header: rt_cuda.h -> librt_cuda.so, rt_cuda_multi.h -> librt_cuda_multi.so (NCCL),
    rt_scene.h -> librt_scene.so (host only: the reference arm of bench.py maps no CUDA library)."""
    total = 0
    for header, path, table in (("rt_cuda.h", pkg.LIB_PATH, pkg.C_ABI), ("rt_scene.h", pkg.SCENE_LIB_PATH, pkg.SCENE_ABI),
                                ("rt_cuda_multi.h", pkg.MULTI_LIB_PATH, pkg.MULTI_ABI)):
        pkg.load()                                  # librt_cuda_multi.so resolves librt_cuda.so through its RPATH
        lib = ctypes.CDLL(str(path))
        names = _declared(header)
        if header == "rt_cuda.h":
            names = [n for n in names if not n.startswith("rt_cuda_multi_")]
        total += len(names)
        for n in names:
            assert hasattr(lib, n), f"{n} declared in include/{header} but not exported by {path.name}"
        bound = {n for n, _, _ in table}            # and the binding covers them all
        assert set(names) <= bound, set(names) - bound
    assert total >= 50
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg.SCENE_LIB_PATH)], capture_output=True, text=True).stdout
    assert "cuda" not in out.lower()
    needed = subprocess.run(["readelf", "-d", str(pkg.SCENE_LIB_PATH)], capture_output=True, text=True).stdout
    assert "cuda" not in needed.lower() and "nccl" not in needed.lower()


def test_multi_layout_matches_the_python_statement(pkg):
    """rt_cuda_multi_shard_rows / _shard_pitch / _locate_row (C) against parallel.py and numpy."""
    import importlib
    par = importlib.import_module(pkg.__name__ + ".parallel")
    for H, W, strip, G in [(45, 70, 4, 2), (101, 33, 16, 8), (7, 5, 3, 4), (64, 64, 16, 4), (4320, 7680, 4, 8), (1, 9, 4, 3)]:
        lay = pkg.multi_layout(W, H, strip, G)
        ref = par.shard_layout(H, W, strip, G)
        assert lay["rows"] == ref["rows"] and lay["pitch"] == ref["pitch"], (H, W, strip, G)
        for row in range(0, H, max(1, H // 23)):
            assert pkg.multi_locate_row(row, strip, G) == par.local_row_of(row, strip, G)


def test_multi_library_refuses_without_a_gpu(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.RtCudaError) as e:
        pkg.MultiRenderer(gpus=2)
    assert e.value.status == -2
    lib = pkg._multi_lib()
    assert lib.rt_cuda_multi_render(None, 8, 8, -4.0, 1.0, 6, 0) == -1
    assert pkg.load().rt_cuda_strerror(-8) == b"RT_CUDA_ERR_NCCL"


def test_signatures_are_plain_c(pkg):
    """extern "C", plain pointers and sizes: no mangled or torch symbols in the dynamic table."""
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg.LIB_PATH)], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(not s.startswith("_Z") or "rtg" in s for s in exported if s.startswith("rt_"))
    assert "rt_cuda_render" in exported and "rt_cuda_readback" in exported
    assert not any("torch" in s or "at::" in s for s in exported)


def test_pod_layouts(pkg):
    assert pkg.SPHERE_DTYPE.itemsize == 48 and pkg.LIGHT_DTYPE.itemsize == 24
    assert pkg.SPHERE_DTYPE.fields["radius"][1] == 12 and pkg.SPHERE_DTYPE.fields["opacity"][1] == 40
    assert ctypes.sizeof(pkg.Stats) % 8 == 0


def test_strerror(pkg):
    lib = pkg.load()
    assert lib.rt_cuda_strerror(0) == b"RT_CUDA_OK"
    assert lib.rt_cuda_strerror(-2) == b"RT_CUDA_ERR_NO_DEVICE"
    assert lib.rt_cuda_strerror(-99) == b"RT_CUDA_ERR_UNKNOWN"


def test_no_cpu_fallback_without_a_gpu(pkg):
    """On a machine without a CUDA device init must fail loudly, never render on the CPU."""
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.RtCudaError) as e:
        pkg.Renderer(0)
    assert e.value.status == -2
    lib = pkg.load()
    assert lib.rt_cuda_render(None, 8, 8, -4.0, 1.0, 6) == -1       # null context: invalid argument
    assert lib.rt_cuda_readback(None, None, None) == -1


def test_product_sources_do_not_touch_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may use oracle/."""
    for p in (ROOT / "raytracer-gamma_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".c", ".cpp"):
            text = p.read_text()
            assert "rt_oracle" not in text and "oracle/" not in text and "hostsim" not in text.replace(
                "tests/hostsim.cpp", ""), p


def test_local_rows_partition(pkg):
    H = 37
    for G in (1, 2, 3, 4, 8):
        rows = [pkg.local_rows(H, 4, g, G) for g in range(G)]
        allr = np.sort(np.concatenate(rows))
        assert np.array_equal(allr, np.arange(H))


def test_scene_file_round_trip(pkg, tmp_path):
    """Hex-float scene files reproduce the arrays bit for bit (incl. awkward values)."""
    sph, lgt = pkg.synth_scene(300, 4, seed=9)
    sph["pos"][3] = (np.float32(1e-40), -0.0, np.float32(3.4e38))     # denormal, -0, near FLT_MAX
    p = tmp_path / "scene.txt"
    pkg.save_scene(p, sph, lgt)
    s2, l2 = pkg.load_scene(p)
    assert np.array_equal(sph.view(np.uint32), s2.view(np.uint32)) and np.array_equal(lgt.view(np.uint32), l2.view(np.uint32))
    text = p.read_text()
    assert text.startswith("rtgamma-scene 1") and text.count("\nsphere ") == 300 and text.count("\nlight ") == 4
    # comments / blank lines are skipped, malformed files are refused
    p.write_text("rtgamma-scene 1\n\n# a comment\nlight 0x1p+0 0 0  1 1 1   # trailing comment\n")
    s3, l3 = pkg.load_scene(p)
    assert len(s3) == 0 and len(l3) == 1 and l3["pos"][0][0] == 1.0
    for bad in ("", "rtgamma-scene 2\n", "rtgamma-scene 1\nsphere 1 2 3\n", "rtgamma-scene 1\ntriangle 1 2 3\n"):
        p.write_text(bad)
        with pytest.raises(OSError):
            pkg.load_scene(p)
    with pytest.raises(OSError):
        pkg.load_scene(tmp_path / "missing.txt")


def test_stats_struct_matches_the_header(pkg, tmp_path):
    """The ctypes mirror of rt_cuda_stats has the C struct's size and field offsets."""
    fields = [n for n, _ in pkg.Stats._fields_]
    src = tmp_path / "sz.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rt_cuda.h"', 'int main(void) {',
             '  printf("%zu\\n", sizeof(rt_cuda_stats));']
    lines += [f'  printf("%zu\\n", offsetof(rt_cuda_stats, {n}));' for n in fields]
    lines += ['  return 0; }']
    src.write_text("\n".join(lines))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(pkg.Stats)
    assert out[1:] == [getattr(pkg.Stats, n).offset for n in fields]


def test_reference_side_program_compiles_against_the_reference_headers(pkg):
    """INTEGRATION.md section 2 as a program (oracle/ref_dropin.cpp): the reference's own structs and
    setters, pointer casts, the C-ABI.  Here it must build and link; without a GPU it must fail loudly."""
    import sys
    graft = sys.modules["__graft_entry__"]
    if not Path("/root/reference/raytracer_gamma/raytracer.h").exists():
        pytest.skip("the reference tree is not on this machine")
    exe = graft.build_dropin()
    assert exe.exists()
    if pkg.device_count() == 0:
        res = subprocess.run([str(exe), "/dev/null"], capture_output=True, text=True, timeout=60)
        assert res.returncode != 0 and "RT_CUDA_ERR_NO_DEVICE" in res.stdout


def test_png_writer_round_trip(tmp_path):
    """host/rt_png.h (the host program's `--out x.png`): signature, chunk CRCs, zlib stream and pixels
    survive a decode with Python's zlib."""
    import struct
    import zlib
    W, H = 37, 23
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    big = rng.integers(0, 256, (200, 150, 3), dtype=np.uint8)        # > 65535 raw bytes: several stored blocks
    src = tmp_path / "png.cpp"
    src.write_text('#include "rt_png.h"\n#include <stdlib.h>\nint main(int c, char** v) { unsigned w = atoi(v[3]), h = atoi(v[4]);\n'
                   '  std::vector<unsigned char> b((size_t)w * h * 3); FILE* f = fopen(v[1], "rb");\n'
                   '  if (!f || fread(b.data(), 1, b.size(), f) != b.size()) return 2; fclose(f);\n'
                   '  return rtpng::write_rgb8(v[2], b.data(), w, h) ? 0 : 1; }\n')
    exe = tmp_path / "png"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(ROOT / "raytracer-gamma_b200" / "host"), "-o", str(exe), str(src)], check=True)
    for arr in (img, big):
        raw, out = tmp_path / "in.rgb", tmp_path / "out.png"
        raw.write_bytes(arr.tobytes())
        subprocess.run([str(exe), str(raw), str(out), str(arr.shape[1]), str(arr.shape[0])], check=True)
        data = out.read_bytes()
        assert data[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(data):
            n, typ = struct.unpack(">I4s", data[pos:pos + 8])
            body = data[pos + 8:pos + 8 + n]
            (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
            assert crc == zlib.crc32(typ + body) & 0xFFFFFFFF
            chunks.append((typ, body))
            pos += 12 + n
        assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        w, h, depth, colour, comp, flt, lace = struct.unpack(">IIBBBBB", chunks[0][1])
        assert (w, h, depth, colour, comp, flt, lace) == (arr.shape[1], arr.shape[0], 8, 2, 0, 0, 0)
        rows = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, w * 3 + 1)
        assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(h, w, 3), arr)
