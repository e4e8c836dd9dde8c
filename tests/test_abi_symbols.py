"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares
(no compute calls here -- those need a GPU), and fails loudly instead of falling back to a CPU path."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()


def test_header_symbols_are_exported():
    _ensure_built()
    from softrender_b200 import _abi
    header = open(os.path.join(ROOT, "include", "softrender_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # declarations only, not prose
    declared = set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(_abi.lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert declared == set(_abi.SYMBOLS), f"binding/header mismatch: {declared ^ set(_abi.SYMBOLS)}"


def test_enum_values_match_header():
    import softrender_b200 as sr
    types = open(os.path.join(ROOT, "include", "softrender_b200_types.h")).read()
    for name, value in re.findall(r"\bSR_([A-Z0-9_]+)\s*=\s*(\d+)", types):
        if hasattr(sr, name):
            assert getattr(sr, name) == int(value), name


def test_uniforms_struct_layout():
    import ctypes
    from softrender_b200.scenes import Light, Uniforms, Viewport
    assert ctypes.sizeof(Light) == 32 and ctypes.sizeof(Viewport) == 24
    assert ctypes.sizeof(Uniforms) == 4 * (4 + 64 + 4 + 4 + 1 + 3) + 8 * 32
    assert Uniforms.lights.offset == 320


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device every entry point must fail loudly (status != 0), never compute on the CPU."""
    _ensure_built()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from softrender_b200 import pipeline
    from softrender_b200._abi import SoftrenderError
    with pytest.raises(SoftrenderError) as e:
        pipeline.Context(0)
    assert e.value.status == 3  # SR_ERR_CUDA


def test_product_sources_do_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "rust-softrender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "sr_oracle" not in text and "oracle_binding" not in text and "libsr_oracle" not in text, f


def test_png_writer_round_trip(tmp_path):
    """`image.save` of examples/suzanne.rs:191 in the host mirror: an 8-bit RGBA PNG that decodes back to the same bytes
    (zlib + CRCs checked chunk by chunk; no GPU involved)."""
    _ensure_built()
    import struct
    import zlib
    import numpy as np
    from softrender_b200 import pipeline
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    path = tmp_path / "t.png"
    pipeline.write_png(str(path), img)
    blob = path.read_bytes()
    assert blob[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, seen = 8, b"", []
    while pos < len(blob):
        n, tag = struct.unpack(">I", blob[pos:pos + 4])[0], blob[pos + 4:pos + 8]
        assert zlib.crc32(blob[pos + 4:pos + 8 + n]) & 0xFFFFFFFF == struct.unpack(">I", blob[pos + 8 + n:pos + 12 + n])[0]
        seen.append(tag)
        if tag == b"IHDR":
            assert struct.unpack(">IIBBBBB", blob[pos + 8:pos + 8 + n]) == (53, 37, 8, 6, 0, 0, 0)
        if tag == b"IDAT":
            idat += blob[pos + 8:pos + 8 + n]
        pos += 12 + n
    assert seen[0] == b"IHDR" and seen[-1] == b"IEND"
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(37, 1 + 53 * 4)
    assert not rows[:, 0].any() and np.array_equal(rows[:, 1:].reshape(37, 53, 4), img)


def test_shader_registry_enumerates_without_a_device():
    """sr_registry_entry: the closures of the reference as a registered set -- ids match the header's enums, layouts match
    the scene definitions (Vin / K sizes of SURVEY section 8), enumeration ends with an error status."""
    _ensure_built()
    import softrender_b200 as sr
    from softrender_b200 import pipeline
    vs, gs, fs, bl = (pipeline.registry(k) for k in range(4))
    assert [e["id"] for e in vs] == [sr.VS_PASSTHROUGH, sr.VS_SUZANNE, sr.VS_FULL_EXAMPLE]
    assert {e["name"]: (e["vin_floats"], e["nk"]) for e in vs}["suzanne"] == (6, 8)        # pos3+normal3 -> {position4, normal4}
    assert {e["name"]: (e["vin_floats"], e["nk"]) for e in vs}["full_example"] == (8, 10)  # + uv2
    assert [e["id"] for e in gs] == [sr.GS_CLIP, sr.GS_FACE_NORMALS, sr.GS_VERTEX_NORMALS, sr.GS_CLIP_SH]
    assert [e["id"] for e in fs] == [sr.FS_FLAT, sr.FS_SUZANNE, sr.FS_FULL_EXAMPLE, sr.FS_FULL_EXAMPLE_TEXTURED, sr.FS_GREEN, sr.FS_DISCARD_CHECKER, sr.FS_TEXTURE_UNLIT,
                                   sr.FS_SUZANNE_GBUFFER]
    assert [e["name"] for e in fs if e["discards"]] == ["discard_checker"] and [e["name"] for e in fs if e["needs_texture"]] == ["full_example_4light_textured", "texture_unlit"]
    assert [e["name"] for e in bl] == ["replace", "alpha_over", "additive"] and pipeline.registry(9) == []
    assert all(e["reference"] for e in vs + gs + fs + bl)


def test_destroy_and_query_entries_accept_null():
    """Every *_destroy accepts a null handle (free(NULL) convention) and the null-argument paths of a few entries return an
    error status instead of dereferencing -- checked without a device."""
    _ensure_built()
    from softrender_b200._abi import lib
    for name in ("sr_context_destroy", "sr_framebuffer_destroy", "sr_mesh_destroy", "sr_texture_destroy", "sr_pipeline_destroy",
                 "sr_draw_destroy"):
        assert getattr(lib, name)(None) == 0, name
    assert lib.sr_context_synchronize(None) != 0
    assert lib.sr_registry_entry(0, 0, None) != 0
    assert lib.sr_last_error()


def test_every_entry_point_survives_null_arguments():
    """Calling each exported function with null handles / pointers and zero scalars returns (an error status for the ones that
    need an object) instead of dereferencing: the boundary never aborts (header conventions)."""
    _ensure_built()
    import ctypes
    from softrender_b200 import _abi
    for name, (_res, args) in sorted(_abi.SYMBOLS.items()):
        vals = [None if (a is ctypes.c_void_p or (isinstance(a, type) and issubclass(a, ctypes._Pointer))) else 0 for a in args]
        getattr(_abi.lib, name)(*vals)  # must return
