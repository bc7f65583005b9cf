"""CPU tests: pin the oracle (oracle/awry_oracle.c) to every literal vector and property the
reference's own tests hold for the search path, plus the Appendix-A worked example.

The reference is Rust and cannot run here, and it ships no data fixtures: its tests are
properties (index == brute force).  They are restated below with the same sizes.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import brute_positions, oracle_from_parts

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "appendix_a.awry")
GOLDEN_SHA256 = "bb57bc33cfedac88a1ccbd980b8af3263be80d1bbb1d4954fb836aa0c26c5f3e"


# ---- literal vectors -------------------------------------------------------------------------

def test_bits_per_element_kat(po):
    """compressed_suffix_array.rs:184-201"""
    L = po.lib()
    for length, bits in [(15, 4), (16, 4), (17, 5), (31, 5), (32, 5), (33, 6), (1022, 10), (1023, 10),
                         (1024, 10), (1025, 11), (65535, 16), (65536, 16), (65537, 17),
                         (2**31 - 1, 31), (2**31, 31), (2**31 + 1, 32)]:
        assert L.awo_bits_per_element(length) == bits, length


def test_alphabet_tables(po):
    """alphabet.rs:169-330 and the round trips of :432-482"""
    L = po.lib()
    dna = {"$": (0, 0b100), "A": (1, 0b110), "C": (2, 0b101), "G": (3, 0b011), "N": (4, 0b010), "T": (5, 0b001)}
    for ch, (idx, code) in dna.items():
        assert L.awo_ascii_to_index(0, ord(ch)) == idx
        assert L.awo_ascii_to_index(0, ord(ch.lower())) == idx
        assert L.awo_index_to_code(0, idx) == code
        assert L.awo_code_to_index(0, code) == idx
        assert L.awo_index_to_ascii(0, idx) == ch.encode()
    assert L.awo_ascii_to_index(0, ord("U")) == 5 and L.awo_ascii_to_index(0, ord("#")) == 0
    assert all(L.awo_ascii_to_index(0, ord(c)) == 4 for c in "RYKMSWBDHVX*-")
    assert L.awo_code_to_index(0, 0b000) == 4 and L.awo_code_to_index(0, 0b111) == 4
    amino = dict(zip("$ACDEFGHIKLMNPQRSTVWXY",
                     [0b00000, 0b01100, 0b10111, 0b00011, 0b00110, 0b11110, 0b11010, 0b11011, 0b11001,
                      0b10101, 0b11100, 0b11101, 0b01000, 0b01001, 0b00100, 0b10011, 0b01010, 0b00101,
                      0b10110, 0b00001, 0b11111, 0b00010]))
    for idx, (ch, code) in enumerate(amino.items()):
        assert L.awo_ascii_to_index(1, ord(ch)) == idx
        assert L.awo_ascii_to_index(1, ord(ch.lower())) == idx
        assert L.awo_index_to_code(1, idx) == code
        assert L.awo_code_to_index(1, code) == idx
        assert L.awo_index_to_ascii(1, idx) == ch.encode()
    assert all(L.awo_ascii_to_index(1, ord(c)) == 20 for c in "BJOUZ*")
    unused = set(range(32)) - set(amino.values())
    assert len(unused) == 10 and all(L.awo_code_to_index(1, c) == 20 for c in unused)


def test_golden_file_is_pinned():
    assert hashlib.sha256(open(GOLDEN, "rb").read()).hexdigest() == GOLDEN_SHA256
    assert os.path.getsize(GOLDEN) == 557


def test_appendix_a_worked_example(po, fx, tmp_path):
    ix = po.OracleIndex.load(GOLDEN)
    assert (ix.bwt_len, ix.sa_ratio, ix.kmer_len, ix.alphabet, ix.bits) == (20, 4, 2, 0, 5)
    assert ix.prefix_sums == [0, 1, 8, 11, 14, 15, 20]
    bwt = "".join(chr(ord(po.lib().awo_index_to_ascii(0, ix.symbol_at(i)))) for i in range(20))
    assert bwt == "TTTNCCGGAAA$ACAGTTAA"
    assert [ix.sa_reconstruct(r) for r in (0, 4, 8, 12, 16)] == [19, 6, 5, 7, 3]
    table = {"A": ((1, 7), [4, 11, 15, 6, 13, 1, 8]), "TTA": ((18, 19), [2, 9]),
             "GATTACA": ((11, 12), [0, 7]), "ACA": ((1, 2), [4, 11]), "N": ((14, 14), [14]),
             "CAN": ((9, 9), [12]), "ACGT": ((3, 3), [15]), "GG": ((13, 12), []),
             "TACAGATTACANACGT": ((16, 16), [3])}
    for q, (rng, locs) in table.items():
        assert ix.search_range(q) == rng, q
        assert ix.count_string(q) == len(locs)
        assert [h[1] for h in ix.locate_string(q)] == locs, q
    # the fixture writer reproduces the pinned bytes (incl. the reference-style k-mer table, Q2)
    parts = fx.build_parts(b"GATTACAGATTACANACGT", 0, ratio=4, kmer_len=2, keep_sa=True)
    assert list(parts.sa) == [19, 4, 11, 15, 6, 13, 1, 8, 5, 12, 16, 0, 7, 17, 14, 18, 3, 10, 2, 9]
    out = tmp_path / "a.awry"
    parts.write(str(out))
    assert hashlib.sha256(out.read_bytes()).hexdigest() == GOLDEN_SHA256


# ---- rank / block properties (bwt.rs:368-505) ---------------------------------------------------

def _mock_index(po, alphabet):
    # a 1-block index just to carry the alphabet shape into awo_block_occurrence
    n_words = 20 if alphabet == 0 else 44
    card = 6 if alphabet == 0 else 22
    return po.OracleIndex.from_parts(alphabet, 8, 200, 4, np.zeros(n_words, np.uint64),
                                     np.zeros(card + 1, np.uint64), np.zeros(64, np.uint64))


@pytest.mark.parametrize("alphabet,seed", [(0, 2), (1, 6)])
def test_random_block_rank_equals_brute_force(po, alphabet, seed):
    """bwt.rs:391-434 / :461-505: inclusive rank == milestone + #sym in 0..=pos, milestones 1000*i"""
    L = po.lib()
    ix = _mock_index(po, alphabet)
    planes, nms, card = (3, 8, 6) if alphabet == 0 else (5, 24, 22)
    rng = np.random.default_rng(seed)
    syms = rng.integers(1, card, 256)
    block = np.zeros(planes * 4 + nms, dtype=np.uint64)
    for pos, s in enumerate(syms):
        code = L.awo_index_to_code(alphabet, int(s))
        for p in range(planes):
            if (code >> p) & 1:
                block[4 * p + pos // 64] |= np.uint64(1) << np.uint64(pos % 64)
    for s in range(card):
        block[4 * planes + s] = 1000 * (s + 1)
    for s in range(1, card):
        run = 0
        for pos in range(256):
            run += int(syms[pos] == s)
            got = L.awo_block_occurrence(ix._h, block.ctypes.data, pos, s)
            assert got == 1000 * (s + 1) + run, (s, pos)
    # empty block => milestone for every position (bwt.rs:368-389 / :436-459); for amino the
    # all-zero code is the sentinel, for DNA 000 matches no predicate
    block[: planes * 4] = 0
    for s in range(1, card):
        for pos in (0, 1, 63, 64, 127, 128, 255):
            assert L.awo_block_occurrence(ix._h, block.ctypes.data, pos, s) == 1000 * (s + 1)


def test_predicates_select_exactly_one_valid_code(po):
    """every plane predicate equals an exact code match on valid codes"""
    L = po.lib()
    for alphabet, planes, card in ((0, 3, 6), (1, 5, 22)):
        ix = _mock_index(po, alphabet)
        valid = [L.awo_index_to_code(alphabet, i) for i in range(card)]
        block = np.zeros(planes * 4 + (8 if alphabet == 0 else 24), dtype=np.uint64)
        for pos, code in enumerate(valid):
            for p in range(planes):
                if (code >> p) & 1:
                    block[4 * p] |= np.uint64(1) << np.uint64(pos)
        for s in range(1, card):
            occ = [L.awo_block_occurrence(ix._h, block.ctypes.data, pos, s) for pos in range(card)]
            marks = [occ[0]] + [occ[i] - occ[i - 1] for i in range(1, card)]
            assert marks == [1 if i == s else 0 for i in range(card)], (alphabet, s)


def test_masked_popcount(po):
    """simd_instructions.rs:96-121: bits 0..=pos inclusive"""
    L = po.lib()
    rng = np.random.default_rng(1)
    v = rng.integers(0, 2**63, 4, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, 4, dtype=np.uint64)
    bits = "".join(format(int(w), "064b")[::-1] for w in v)
    for pos in range(256):
        assert L.awo_masked_popcount(v.ctypes.data, pos) == bits[: pos + 1].count("1")


# ---- sampled suffix array (compressed_suffix_array.rs:137-180) ------------------------------------

def test_sampled_sa_round_trip(po):
    L = po.lib()
    sa_len = 123451
    rng = np.random.default_rng(0)
    for ratio in range(1, 16):
        bits = L.awo_bits_per_element(sa_len)
        n_words = L.awo_compressed_word_len(sa_len, ratio)
        words = np.zeros(n_words + 1, dtype=np.uint64)
        n_el = -(-sa_len // ratio)
        vals = rng.integers(0, sa_len, n_el)
        for e, v in enumerate(vals):
            L.awo_sa_set_value(words.ctypes.data, bits, int(v), e)
        ix = po.OracleIndex.from_parts(0, ratio, sa_len, 4, np.zeros(20 * (-(-sa_len // 256)), np.uint64),
                                       np.zeros(7, np.uint64), words[:n_words])
        for e in list(range(0, n_el, 997)) + [n_el - 1]:
            assert ix.sa_reconstruct(e * ratio) == int(vals[e])


# ---- end-to-end vs brute force (fm_index.rs:612-743) ------------------------------------------------

def _all_kmers(text: bytes, k: int):
    seen = {}
    for p in range(0, len(text) - k + 1):
        seen.setdefault(text[p:p + k], []).append(p)
    return seen


def test_nucleotide_index_every_24mer(po, fx):
    """fm_index.rs:666-700: 1847-bp random FASTA, defaults ratio 8 / k 10, every distinct 24-mer"""
    text = fx.gen_text(0, 1847, 0)
    parts = fx.build_parts(text, 0)          # defaults: ratio 8, k 10 (kmer_lookup_table.rs:23)
    ix = oracle_from_parts(po, parts)
    kmers = _all_kmers(bytes(text), 24)
    qb, qo = po.pack_queries(list(kmers))
    counts, _ = ix.count_batch(qb, qo)
    off, hits, _ = ix.locate_batch(qb, qo, sorted_hits=True)
    for i, (q, positions) in enumerate(kmers.items()):
        assert int(counts[i]) == len(positions)
        assert [int(x) for x in hits[int(off[i]):int(off[i + 1]), 1]] == positions
        assert ix.count_string(q) == len(positions)


def test_amino_index_every_8mer(po, fx):
    """fm_index.rs:702-743: 300 residues, 8-mers, defaults ratio 8 / k 4"""
    text = fx.gen_text(1, 300, 999)
    parts = fx.build_parts(text, 1)
    ix = oracle_from_parts(po, parts)
    for q, positions in _all_kmers(bytes(text), 8).items():
        assert ix.count_string(q) == len(positions)
        assert sorted(h[1] for h in ix.locate_string(q)) == positions


@pytest.mark.parametrize("ratio", [1, 3, 8, 32])
def test_random_queries_vs_brute_force(po, fx, ratio):
    text = fx.gen_text(0, 5000, 40 + ratio)
    parts = fx.build_parts(text, 0, ratio=ratio, kmer_len=5)
    ix = oracle_from_parts(po, parts)
    t = bytes(text)
    rng = np.random.default_rng(ratio)
    for _ in range(300):
        n = int(rng.integers(1, 12))
        p = int(rng.integers(0, len(t) - n))
        q = t[p:p + n] if rng.random() < 0.7 else bytes(rng.choice(list(b"ACGT"), n).tolist())
        want = brute_positions(t, q)
        assert ix.count_string(q) == len(want), q
        assert sorted(h[1] for h in ix.locate_string(q)) == want, q


def test_multi_record_every_suffix_is_found(po, fx):
    """fm_index.rs:779-790, :1034-1045 (count only, as in the reference) + intended locate mapping"""
    rng = np.random.default_rng(5)
    recs = [bytes(rng.choice(list(b"ACGT"), int(n)).tolist()) for n in rng.integers(5, 59, 30)]
    text, starts = fx.concat_records(recs, 0)
    parts = fx.build_parts(text, 0, seq_starts=starts, kmer_len=4)
    ix = oracle_from_parts(po, parts)
    for ri, r in enumerate(recs):
        for s in range(len(r)):
            assert ix.count_string(r[s:]) != 0
        hits = ix.locate_string(r)
        assert (ri, 0) in hits


def test_save_load_round_trip(po, fx, tmp_path):
    """fm_index.rs:1046-1088: field-by-field equality after save -> load"""
    text = fx.gen_text(0, 3000, 8)
    parts = fx.build_parts(text, 0, ratio=5, kmer_len=3)
    path = str(tmp_path / "rt.awry")
    parts.write(path)
    ix = po.OracleIndex.load(path)
    assert (ix.bwt_len, ix.sa_ratio, ix.kmer_len, ix.alphabet, ix.version) == (3001, 5, 3, 0, 1)
    assert ix.prefix_sums == [int(x) for x in parts.prefix_sums]
    ref = oracle_from_parts(po, parts)
    for row in range(0, 3001, 7):
        assert ix.symbol_at(row) == ref.symbol_at(row)
        assert ix.backstep(row) == ref.backstep(row)
    for row in range(0, 3001, 5):
        assert ix.sa_reconstruct(row) == ref.sa_reconstruct(row)
    expected_size = 11 + 32 + 12 * 160 + 56 + 8 * len(parts.sa_words) + 1 + 16 * 4**3 + 8 + 16 + 9
    assert os.path.getsize(path) == expected_size


# ---- semantics of the search loop ---------------------------------------------------------------

def test_interval_semantics(po, fx):
    """search.rs:88-144 + fm_index.rs:402-438: both arms (len < k, len >= k) agree with the plain
    backward search; empty intervals stay empty; errors where the reference panics"""
    text = fx.gen_text(0, 4000, 3)
    t = bytes(text)
    a = oracle_from_parts(po, fx.build_parts(text, 0, kmer_len=2))
    b = oracle_from_parts(po, fx.build_parts(text, 0, kmer_len=12))
    for q in [t[10:11], t[10:14], t[10:21], t[10:22], t[10:23], t[100:160], b"ACGTACGTACGTAAAA", b"N", b"GNA"]:
        ca, cb = a.count_string(q), b.count_string(q)
        assert ca == cb == len(brute_positions(t, q)), q
    sp, ep = a.search_range(b"ACGTACGTACGTACGTACGTAAAATTTT")
    assert sp > ep
    assert a.update_range(5, 4, 1)[0] == a.update_range(5, 4, 1)[1] + 1   # empty stays empty
    for bad in (b"", b"AC$", b"#"):
        with pytest.raises(po.OracleError):
            a.count_string(bad)
    # lowercase / U / unknown letters are searched as N (alphabet.rs:169-248)
    assert a.count_string(t[50:70].lower()) == a.count_string(t[50:70])
    assert a.count_string(t[50:70].replace(b"T", b"U")) == a.count_string(t[50:70])
    assert a.count_string(b"R") == a.count_string(b"N") == 0


def test_batch_equals_single_and_stats(po, fx):
    text = fx.gen_text(0, 20000, 13)
    parts = fx.build_parts(text, 0, kmer_len=6)
    ix = oracle_from_parts(po, parts)
    qb, qo, _ = fx.gen_substring_queries(text, 500, 30, 2)
    for threads in (1, 3):
        counts, st = ix.count_batch(qb, qo, n_threads=threads)
        assert all(int(counts[i]) == ix.count_string(bytes(qb[30 * i:30 * i + 30])) for i in range(0, 500, 17))
        assert st["lf_steps"] == 500 * 29 and st["seeded_steps"] == 500 * 24
        assert st["seeded_steps"] <= st["seeded_touches"] <= 2 * st["seeded_steps"]
    off, hits, st = ix.locate_batch(qb, qo, n_threads=2)
    assert st["hits"] == len(hits) == int(off[-1])
    with pytest.raises(po.OracleError):
        ix.count_batch(*po.pack_queries([b"ACGT", b""]))
