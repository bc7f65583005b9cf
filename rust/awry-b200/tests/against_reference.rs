//! The one-command pin against the REAL reference crate (`awry` 0.3.1) for a maintainer with a Rust toolchain.
//!
//! This repository was developed without cargo/rustc and without the crate registry, so the CUDA path is
//! pinned by the reference's own test *properties* (index answers == brute force over the text,
//! fm_index.rs:612-743) through a C restatement of its algorithm (`oracle/`), not by outputs of the reference
//! binary.  This test closes that gap wherever both crates can run: it builds the same seeded FASTA with
//! both, and compares
//!   1. `parallel_count`           (fm_index.rs:455-460)  -- every count, input order;
//!   2. `parallel_locate`          (fm_index.rs:479-487)  -- every hit list, in the reference's own push
//!                                                           order (BWT rows ascending, fm_index.rs:521);
//!   3. `save` files               (fm_index_file.rs:42-106) -- byte for byte, both directions: the file the
//!      reference wrote loads here and re-saves identically; the file written here loads in the reference
//!      and compares equal field by field (the reference's own save/load test, fm_index.rs:1046-1088).
//! Single-record inputs only for locate (multi-record locate does not terminate in the reference,
//! sequence_index.rs:108-141).
//!
//! Run:  scripts/pin_against_reference.sh     (needs cargo, a B200, and `awry = "0.3.1"` resolvable)
//! `awry` is a dev-dependency added by that script, so that the crate builds without the registry otherwise.
#![cfg(feature = "against-reference")]

use rayon::prelude::*;
use std::io::Write;
use std::path::{Path, PathBuf};

// the seeded generator of fixtures/fixture_cpu.cpp (splitmix-style), so that this test, the Python tests and
// bench.py search the same synthetic texts
fn mix64(mut z: u64) -> u64 {
    z = z.wrapping_add(0x9E3779B97F4A7C15);
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^ (z >> 31)
}
fn rnd_at(seed: u64, i: u64) -> u64 {
    mix64(mix64(seed) ^ i.wrapping_mul(0xD1342543DE82EF95))
}
fn synth_text(amino: bool, n: u64, seed: u64) -> Vec<u8> {
    const AA: &[u8; 20] = b"ACDEFGHIKLMNPQRSTVWY";
    (0..n)
        .map(|i| {
            let r = rnd_at(seed, i);
            if amino {
                AA[((((r >> 32) * 20) >> 32) as usize)]
            } else {
                b"ACGT"[(r >> 62) as usize]
            }
        })
        .collect()
}
fn write_fasta(path: &Path, text: &[u8]) {
    let mut f = std::fs::File::create(path).unwrap();
    writeln!(f, ">synthetic").unwrap();
    for line in text.chunks(80) {
        f.write_all(line).unwrap();
        f.write_all(b"\n").unwrap();
    }
}
fn queries(text: &[u8], nq: usize, len: usize, seed: u64) -> Vec<String> {
    // half exact substrings, half random strings of the same alphabet (BASELINE cfg1), plus edge cases
    let mut out = Vec::new();
    let span = (text.len() - len + 1) as u64;
    for q in 0..nq as u64 {
        if q % 2 == 0 {
            let p = ((rnd_at(seed, q) as u128 * span as u128) >> 64) as usize;
            out.push(String::from_utf8(text[p..p + len].to_vec()).unwrap());
        } else {
            let alphabet: Vec<u8> = { let mut a = text[..200.min(text.len())].to_vec(); a.sort(); a.dedup(); a };
            out.push((0..len).map(|j| alphabet[(rnd_at(seed ^ 77, q * 1000 + j as u64) % alphabet.len() as u64) as usize] as char).collect());
        }
    }
    for n in 1..14 { out.push(String::from_utf8(text[5..5 + n].to_vec()).unwrap()); }   // shorter than / equal to / above k
    out.push(String::from_utf8(text[40..72].to_vec()).unwrap().to_lowercase());
    out.push("N".into());
    out
}

fn compare(amino: bool, n: u64, text_seed: u64, ratio: u64, k: u8, qlen: usize, dir: &Path) {
    let text = synth_text(amino, n, text_seed);
    let fasta = dir.join(format!("t{}_{}.fa", amino as u8, n));
    write_fasta(&fasta, &text);
    let qs = queries(&text, 4000, qlen, 2);

    // ---- the reference
    let ref_args = awry::fm_index::FmBuildArgs {
        input_file_src: fasta.clone(),
        suffix_array_output_src: Some(dir.join("ref.sufr")),
        suffix_array_compression_ratio: Some(ratio),
        lookup_table_kmer_len: Some(k),
        alphabet: if amino { awry::alphabet::SymbolAlphabet::Amino } else { awry::alphabet::SymbolAlphabet::Nucleotide },
        max_query_len: None,
        remove_intermediate_suffix_array_file: true,
    };
    let reference = awry::fm_index::FmIndex::new(&ref_args).expect("reference build");
    let ref_file = dir.join("ref.awry");
    reference.save(&ref_file).expect("reference save");
    let want_counts = reference.parallel_count(qs.par_iter().map(|s| s.as_str()));
    let want_hits = reference.parallel_locate(qs.par_iter().map(|s| s.as_str()));

    // ---- this crate, (a) loading the reference's file, (b) building on the GPU
    let loaded = awry_b200::fm_index::FmIndex::load(&ref_file).expect("load of the reference's file");
    let built = awry_b200::fm_index::FmIndex::new(&awry_b200::fm_index::FmBuildArgs::new(
        fasta.clone(), None, Some(ratio), Some(k),
        if amino { awry_b200::alphabet::SymbolAlphabet::Amino } else { awry_b200::alphabet::SymbolAlphabet::Nucleotide },
        None, true)).expect("GPU build");
    for (name, ix) in [("loaded", &loaded), ("built", &built)] {
        assert_eq!(ix.bwt_len(), reference.bwt_len(), "{name}");
        assert_eq!(ix.prefix_sums(), reference.prefix_sums(), "{name}");
        let counts = ix.parallel_count(qs.par_iter().map(|s| s.as_str()));
        assert_eq!(counts, want_counts, "{name}: parallel_count");
        let hits = ix.parallel_locate(qs.par_iter().map(|s| s.as_str()));
        assert_eq!(hits.len(), want_hits.len());
        for (q, (got, want)) in hits.iter().zip(want_hits.iter()).enumerate() {
            let g: Vec<(usize, usize)> = got.iter().map(|h| (h.sequence_idx(), h.local_position())).collect();
            let w: Vec<(usize, usize)> = want.iter().map(|h| (h.sequence_idx(), h.local_position())).collect();
            assert_eq!(g, w, "{name}: parallel_locate, query {q} ({})", qs[q]);
        }
        // file interop, both directions
        let ours = dir.join(format!("{name}.awry"));
        ix.save(&ours).expect("save");
        assert_eq!(std::fs::read(&ours).unwrap(), std::fs::read(&ref_file).unwrap(), "{name}: save is not byte-identical");
        let back = awry::fm_index::FmIndex::load(&ours).expect("the reference loads our file");
        assert_eq!(back.parallel_count(qs.par_iter().map(|s| s.as_str())), want_counts);
    }
}

fn scratch() -> PathBuf {
    let d = std::env::temp_dir().join("awry_b200_against_reference");
    std::fs::create_dir_all(&d).unwrap();
    d
}

#[test]
fn nucleotide_1mbp_cfg1() { compare(false, 1_000_000, 1, 8, 10, 32, &scratch()); }

#[test]
fn nucleotide_ratio_and_k_variants() {
    compare(false, 200_000, 5, 1, 4, 20, &scratch());
    compare(false, 200_000, 5, 32, 13, 50, &scratch());
}

#[test]
fn protein_300k_residues() { compare(true, 300_000, 6, 8, 4, 12, &scratch()); }
