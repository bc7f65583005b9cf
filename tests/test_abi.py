"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/awry_b200.h declares, compiles as C and C++, and fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu

HEADER = os.path.join(ROOT, "include", "awry_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(awry_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from awry_b200 import fm_index as f
    lib = f.native()
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(f.EXPORTS) == syms
    assert b"sm_100a" in lib.awry_version()


def test_header_is_plain_c_and_cxx(tmp_path):
    for comp, ext, std in (("gcc", "c", "-std=c99"), ("g++", "cpp", "-std=c++11")):
        src = tmp_path / f"t.{ext}"
        src.write_text('#include "awry_b200.h"\nint main(void){ awry_hit h; h.seq_idx = 0; return (int)h.seq_idx + (int)sizeof(awry_info) * 0; }\n')
        subprocess.check_call([comp, std, "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                               "-c", str(src), "-o", str(tmp_path / f"t_{ext}.o")])


def test_struct_layouts_match_ctypes(tmp_path):
    """ctypes mirrors vs the compiler's view of include/awry_b200.h"""
    from awry_b200 import fm_index as f
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "awry_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(awry_range), sizeof(awry_parts), sizeof(awry_info), sizeof(awry_profile), sizeof(awry_hit), '
                   'sizeof(awry_build_args));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes == [C.sizeof(f._Range), C.sizeof(f._Parts), C.sizeof(f._Info), C.sizeof(f.Profile), 16,
                     C.sizeof(f.BuildArgs)]


def test_kernels_are_sm100a_native():
    """the shipped library carries sm_100a SASS with the 256-bit global loads of the rank kernels"""
    lib = os.path.join(ROOT, "awry_b200", "libawry_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    assert "search_dna_pair_kernel" in sass and "walk_dna_kernel" in sass
    assert re.search(r"LDG\.E\S*\.256", sass) and "POPC" in sass and "SHFL.BFLY" in sass


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly(tmp_path):
    from awry_b200 import AwryError, FmIndex
    golden = os.path.join(ROOT, "tests", "golden", "appendix_a.awry")
    with pytest.raises(AwryError) as e:
        FmIndex.load(golden)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    with pytest.raises(AwryError):
        FmIndex.from_parts(0, 8, 1000, 4, np.zeros(80, np.uint64), np.zeros(7, np.uint64), np.zeros(64, np.uint64))


def test_bad_arguments_are_errors_not_crashes(tmp_path):
    from awry_b200 import fm_index as f
    lib = f.native()
    out = C.c_void_p()
    assert lib.awry_index_load(None, None, 0, C.byref(out)) == -1
    assert lib.awry_index_load(b"/nonexistent/x.awry", None, 0, C.byref(out)) == -2
    assert b"cannot open" in lib.awry_last_error()
    bad = tmp_path / "bad.awry"
    bad.write_bytes(b"not an index at all, just some bytes....................")
    assert lib.awry_index_load(os.fsencode(str(bad)), None, 0, C.byref(out)) == -3
    assert lib.awry_count_batch(None, None, None, 5, None) == -1
    assert lib.awry_index_info(None, None) == -1
    lib.awry_index_free(None)
    lib.awry_hits_free(None)
    assert lib.awry_profile_get(None) == -1
    assert lib.awry_set_search_variant(3, 0, 0) == -1
    n = C.c_uint64()
    assert lib.awry_count_reads_file(None, b"x.fq", C.byref(out), C.byref(n)) == -1
    lib.awry_buffer_free(None)
    # construction: argument and input-file errors are reported before any device is needed
    assert lib.awry_index_build(None, None, 0, C.byref(out)) == -1
    assert lib.awry_build_index_file(None) == -1
    a = f.BuildArgs(b"/nonexistent/in.fa", os.fsencode(str(tmp_path / "o.awry")), 0, 0, 0, 0)
    assert lib.awry_build_index_file(C.byref(a)) == -2 and b"cannot open" in lib.awry_last_error()
    notfa = tmp_path / "reads.txt"
    notfa.write_text("ACGT\n")
    a = f.BuildArgs(os.fsencode(str(notfa)), None, 0, 0, 0, 0)
    assert lib.awry_build_index_file(C.byref(a)) == -1          # no output file
    assert lib.awry_index_build(C.byref(a), None, 0, C.byref(out)) == -2 and b"FASTA" in lib.awry_last_error()
    a = f.BuildArgs(os.fsencode(str(notfa)), None, 7, 0, 0, 0)
    assert lib.awry_index_build(C.byref(a), None, 0, C.byref(out)) == -1
    assert lib.awry_parts_num_blocks(257) == 2 and lib.awry_parts_block_words(0) == 20 and lib.awry_parts_block_words(1) == 44
    assert lib.awry_parts_sa_words(123451, 8) == (15432 * 17 + 63) // 64   # compressed_suffix_array.rs:113-130


def test_truncated_index_files_are_refused_before_any_device_work(tmp_path):
    """the loader checks the sizes its header implies against the file before it seeks or allocates"""
    from awry_b200 import fm_index as f
    lib = f.native()
    out = C.c_void_p()
    golden = open(os.path.join(ROOT, "tests", "golden", "appendix_a.awry"), "rb").read()
    for cut in (12, 43, 100, 203, 250, 270, len(golden) - 40):
        p = tmp_path / f"cut{cut}.awry"
        p.write_bytes(golden[:cut])
        assert lib.awry_index_load(os.fsencode(str(p)), None, 0, C.byref(out)) == -3, cut
    huge = bytearray(golden)
    huge[27:35] = (2**40).to_bytes(8, "little")            # bwt_len far beyond the file
    p = tmp_path / "huge.awry"
    p.write_bytes(bytes(huge))
    assert lib.awry_index_load(os.fsencode(str(p)), None, 0, C.byref(out)) == -3
    assert b"shorter" in lib.awry_last_error()


def test_file_section_offsets_match_the_reference_format():
    """.awry v1 section offsets at the BASELINE sizes (derived from fm_index_file.rs:42-106,
    :165-181 and kmer_lookup_table.rs:55-77), computed with the size helpers the writer and loader use"""
    from awry_b200 import fm_index as f
    lib = f.native()

    def offsets(alphabet, bwt_len, ratio, k):
        card = 6 if alphabet == 0 else 22
        blocks = 43                                            # 11-byte label + 4 x u64 header
        prefix = blocks + lib.awry_parts_num_blocks(bwt_len) * lib.awry_parts_block_words(alphabet) * 8
        sa = prefix + (card + 1) * 8
        kbyte = sa + lib.awry_parts_sa_words(bwt_len, ratio) * 8
        seq = kbyte + 1 + 16 * (card - 2) ** k
        return prefix, sa, kbyte, seq

    assert offsets(0, 3_100_000_001, 8, 13) == (1_937_500_203, 1_937_500_259, 3_487_500_267, 4_561_242_092)
    assert offsets(1, 2_000_000_001, 8, 5) == (2_750_000_395, 2_750_000_579, 3_718_750_587, 3_769_950_588)
    assert offsets(0, 20, 4, 2) == (203, 259, 267, 524)        # Appendix A: 557 bytes = 524 + 8 + 16 + 9
    golden = open(os.path.join(ROOT, "tests", "golden", "appendix_a.awry"), "rb").read()
    assert len(golden) == 557 and golden[267] == 2
    assert int.from_bytes(golden[203 + 48:203 + 56], "little") == 20 and int.from_bytes(golden[524:532], "little") == 1


def test_rust_pin_harness_is_shipped():
    """the one-command check against the real crate (needs a Rust toolchain, absent here) exists and names the
    reference calls it compares"""
    rs = open(os.path.join(ROOT, "rust", "awry-b200", "tests", "against_reference.rs")).read()
    for call in ("awry::fm_index::FmIndex::new", "parallel_count", "parallel_locate", ".save(", "FmIndex::load"):
        assert call in rs, call
    sh = open(os.path.join(ROOT, "scripts", "pin_against_reference.sh")).read()
    assert "cargo test" in sh and "against_reference" in sh


def test_product_never_touches_the_oracle():
    """the shipped package must not import, link or execute anything under oracle/ or fixtures/"""
    pkg = os.path.join(ROOT, "awry_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in txt.replace("oracle/ or fixtures/", ""), (fn, "mentions oracle")
                assert "pyfixture" not in txt and "awo_" not in txt, fn
    ldd = subprocess.run(["ldd", os.path.join(pkg, "libawry_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "fixture" not in ldd


def test_rust_sys_crate_declares_every_symbol():
    """the (uncompilable here) Rust -sys crate must at least name every function of the C header"""
    rs = open(os.path.join(ROOT, "rust", "awry-b200-sys", "src", "lib.rs")).read()
    have = set(re.findall(r"pub fn (awry_[a-z0-9_]+)", rs))
    assert sorted(have) == declared_symbols()


def test_integration_doc_names_every_symbol():
    """INTEGRATION.md is the maintainer's map of the boundary: every entry point of the header appears in it"""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert [s for s in declared_symbols() if s not in doc] == []


def test_every_declaration_cites_the_reference_or_says_why_not():
    """each block of the header names the reference item it replaces (file:line) or is one of the sections that
    have no counterpart (device-resident entry points, instrumentation, knobs)"""
    src = open(HEADER).read()
    cites = re.findall(r"[a-z_]+\.rs:\d+", src)
    assert len(cites) >= 40
    for name in ("awry_index_load", "awry_index_save", "awry_index_from_parts", "awry_count_batch", "awry_locate_batch",
                 "awry_search_batch", "awry_update_range", "awry_backstep", "awry_initial_range", "awry_index_build"):
        at = src.index("int " + name)
        comment = src[src.rfind("/*", 0, at):at]
        assert re.search(r"\.rs:\d+", comment), name


def test_cxx_mirror_compiles_against_the_header():
    """include/awry_b200.hpp (the C++ mirror of the Rust API) and the test program that uses it compile cleanly;
    the program itself runs in the GPU suite (test_cxx_host_mirror_every_kmer)"""
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I",
                           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cxx_host_mirror.cpp")])
