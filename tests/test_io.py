"""Input / output contract of the host pipeline: pair.txt, cam files, .npy writer.  CPU only."""
import ctypes as C

import numpy as np

import capi
import synth


def test_npy_writer_matches_reference_header(tmp_path):
    lib = capi.load()
    for arr, descr in [(np.arange(12, dtype=np.float32).reshape(3, 4), b"<f4"),
                       (np.arange(24, dtype=np.float32).reshape(2, 4, 3), b"<f4"),
                       (np.arange(12, dtype=np.int8).reshape(3, 4), b"|i1")]:
        p = tmp_path / "a.npy"
        ch = arr.shape[2] if arr.ndim == 3 else 1
        assert lib.dpe_host_write_npy(str(p).encode(), arr.ctypes.data_as(C.c_void_p), descr, arr.itemsize, arr.shape[0], arr.shape[1], ch) == 0
        raw = p.read_bytes()
        assert raw[:8] == b"\x93NUMPY\x01\x00"
        hlen = int.from_bytes(raw[8:10], "little")
        assert (10 + hlen) % 16 == 0                       # main.cpp:78-81
        header = raw[10:10 + hlen].decode()
        shape = ", ".join(str(s) for s in arr.shape)
        assert header.startswith("{'descr': '%s', 'fortran_order': False, 'shape': (%s), }" % (descr.decode(), shape))
        assert header.endswith("\n")
        back = np.load(p)
        assert back.dtype == arr.dtype and np.array_equal(back, arr)


def test_cam_and_pair_parsers(tmp_path):
    lib = capi.load()
    K = np.array([[576.0, 0, 320], [0, 577.5, 240], [0, 0, 1]])
    R = np.eye(3)[[1, 0, 2]].astype(float)
    t = np.array([0.1, -0.2, 3.0])
    synth.write_cam(tmp_path / "c.txt", K, R, t, 1.25, 9.5)
    out = np.zeros(23, np.float32)
    assert lib.dpe_host_read_cam(str(tmp_path / "c.txt").encode(), out.ctypes.data_as(C.c_void_p)) == 0
    assert np.allclose(out[:9].reshape(3, 3), K) and np.allclose(out[9:18].reshape(3, 3), R) and np.allclose(out[18:21], t)
    assert abs(out[21] - 1.25) < 1e-6 and abs(out[22] - 9.5) < 1e-6
    Kp, Rp, tp, dmin, dmax = synth.read_cam(tmp_path / "c.txt")
    assert np.allclose(Kp, K) and np.allclose(tp, t) and (dmin, dmax) == (1.25, 9.5)
    # pair.txt: sources with score <= 0 are dropped (main.cpp:301-303); ragged lists; empty list
    (tmp_path / "pair.txt").write_text("3\n7\n3 1 10.5 2 0.0 3 2.0\n1\n0\n2\n2 7 -1 1 5\n")
    buf = np.zeros(64, np.int32)
    n = lib.dpe_host_read_pairs(str(tmp_path / "pair.txt").encode(), buf.ctypes.data_as(C.c_void_p), 64)
    assert list(buf[:n]) == [7, 2, 1, 3, 1, 0, 2, 1, 1]
    assert lib.dpe_host_read_pairs(str(tmp_path / "missing.txt").encode(), buf.ctypes.data_as(C.c_void_p), 64) == -1


def test_malformed_inputs_are_refused_quickly(tmp_path):
    """Files that announce more than they hold.  A count in pair.txt that overflows an int used to become the bound of a
    loop of failed extractions (the reference's GenerateSampleList has the same loop); a .dmb header with huge
    dimensions used to size an allocation before the payload was looked at.  Both are plain errors now — found by running
    the parsers under -fsanitize=address,undefined on a corpus of damaged files."""
    import struct
    import time
    lib = capi.load()
    lib.dpe_host_read_dmb.restype = C.c_long
    buf = np.zeros(64, np.int32)
    t0 = time.perf_counter()
    for text in ("999999999999999999999\n0\n1 1 100.0\n",            # number of problems overflows
                 "2\n0\n999999999999999999999 1 100.0\n1\n1 0 100.0\n",  # number of sources overflows
                 "3\n0\n1 1 100.0\n",                                 # fewer problems than announced
                 "x\n", "", "2\n0\n"):
        (tmp_path / "pair.txt").write_text(text)
        assert lib.dpe_host_read_pairs(str(tmp_path / "pair.txt").encode(), buf.ctypes.data_as(C.c_void_p), 64) == -1, text
    # fewer sources on a line than announced: what is there is kept
    (tmp_path / "pair.txt").write_text("1\n4\n5 1 10.0 2 20.0\n")
    n = lib.dpe_host_read_pairs(str(tmp_path / "pair.txt").encode(), buf.ctypes.data_as(C.c_void_p), 64)
    assert list(buf[:n]) == [4, 2, 1, 2]
    r, c, t = C.c_int(), C.c_int(), C.c_int()
    good = struct.pack("<4i", 1, 3, 5, 0) + bytes(range(15))
    (tmp_path / "g.dmb").write_bytes(good)
    assert lib.dpe_host_read_dmb(str(tmp_path / "g.dmb").encode(), C.byref(r), C.byref(c), C.byref(t)) == 15 and (r.value, c.value, t.value) == (3, 5, 0)
    for bad in (struct.pack("<4i", 1, 2**31 - 1, 2**31 - 1, 21) + b"abc",      # 2^62 elements announced
                struct.pack("<4i", 1, 3, 5, 4) + bytes(59),                      # one byte short
                struct.pack("<4i", 1, 3, 5, 7) + bytes(64),                      # unknown type
                struct.pack("<4i", 2, 3, 5, 0) + bytes(15), good[:10], b""):
        (tmp_path / "b.dmb").write_bytes(bad)
        assert lib.dpe_host_read_dmb(str(tmp_path / "b.dmb").encode(), C.byref(r), C.byref(c), C.byref(t)) == -1
    assert time.perf_counter() - t0 < 5.0


def test_pipeline_refuses_bad_scenes_before_touching_a_gpu(tmp_path, capfd):
    """The checks of GenerateSampleList / CheckImages / InuputInitialization that dpe_run_pipeline makes before it needs a
    device: no pair.txt, a reference image listed twice, more than 31 sources (DPE.cpp:762-765: "Can't process so
    much images") — each returns non-zero with the reference's message, with or without a GPU in the box."""
    lib = capi.load()
    run = lambda: lib.dpe_run_pipeline(str(tmp_path).encode(), 0, 0, 0, 0, 1, 0, 0, 0)
    assert run() != 0
    assert "Images may error" in capfd.readouterr().err
    (tmp_path / "pair.txt").write_text("2\n0\n1 1 100.0\n0\n1 1 100.0\n")
    assert run() != 0
    assert "twice" in capfd.readouterr().err
    (tmp_path / "pair.txt").write_text("1\n0\n32 " + " ".join(f"{i + 1} 100.0" for i in range(32)) + "\n")
    assert run() != 0
    assert "Can't process so much images: 33" in capfd.readouterr().err
