"""bench.py picks the DRAM-traffic figure of each kernel flavour from the committed ncu summaries (profiles/*_traffic.json):
the lookups must find the captures the roofline blocks are computed from, and tell the two flavours of the pair kernel apart."""
import importlib.util
import os

from conftest import ROOT


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_traffic_lookup_finds_every_capture():
    b = load_bench()
    wave = b.traffic_from_profiles("search_dna_wave_kernel")
    lf = b.traffic_from_profiles("search_dna_pair_kernel", in_text=False)
    amino = b.traffic_from_profiles("search_amino")
    assert wave and wave["kernel"].startswith("search_dna_wave_kernel") and wave["dram_bytes_per_launch"] > 1e9
    assert lf and lf["kernel"].startswith("search_dna_pair_kernel") and not lf["kernel"].rstrip().endswith(", 1>")
    assert amino and amino["kernel"].startswith("search_amino") and amino["dram_bytes_per_launch"] > 1e9
    # backward search to the last symbol moves an order of magnitude more than the default path
    assert lf["dram_bytes_per_launch"] > 5 * wave["dram_bytes_per_launch"]
    assert b.traffic_from_profiles("no_such_kernel") is None


def test_both_arms_describe_the_same_config():
    b = load_bench()

    class A:
        reads, read_len, text_len, kmer, sa_ratio = 10_000_000, 150, 3_100_000_000, 13, 8
    cfg = b.config_of(A)
    assert "workload" in cfg and "10000000 x 150-bp" in cfg["workload"] and "model" not in cfg
