"""produce_hpf: per-population `<POP>.freqs.gz` files -> hpf.csv + the population counts file.

Same role, arguments and output files as graph_generation/generate_hpf.py:8-83 of the reference
(the first step of its README example): host-side data-format work that runs once per frequency
set, before `grim.grim.graph_freqs` builds the device tables from hpf.csv.

Input line:  Haplo,Count,Freq  (header `Haplo,...` skipped; rows with frequency 0 dropped).
hpf.csv:     header `hap,pop,freq`, then one row per (population, haplotype) in reading order, a
             later duplicate replacing the earlier value in place; frequencies printed as Python
             prints floats.
counts file: `<pop>,<sum of Count over the kept rows>,<share of the grand total>` per population
             (column 3 is what impute() multiplies the prior matrix with, impute.py:205-212).
Population names must not contain '-' (the reference joins pop and haplotype with it)."""
import argparse
import gzip
import json
import os

project_dir = ""   # prefix of every configured path, as in the reference module


def produce_hpf(conf_file):
    with open(conf_file) as f:
        conf = json.load(f)
    pops = conf.get("populations")
    freq_dir = project_dir + conf.get("freq_data_dir")
    out_dir = project_dir + conf.get("graph_files_path")
    counts_path = project_dir + conf.get("pops_count_file")
    hpf_path = project_dir + conf.get("freq_file")
    os.makedirs(out_dir, exist_ok=True)

    freqs = {}       # (pop, haplotype) -> frequency, insertion ordered
    totals = []
    for pop in pops:
        total = 0
        with gzip.open(os.path.join(freq_dir, pop + ".freqs.gz"), "rt", encoding="utf8") as zf:
            for line in zf:
                haplotype, count, freq = line.strip().split(",")
                if haplotype == "Haplo":
                    continue
                freq = float(freq)
                if freq == 0.0:
                    continue
                freqs[(pop, haplotype)] = freq
                total += float(count)
        totals.append(total)

    grand = sum(totals)
    with open(counts_path, "w") as f:
        for pop, total in zip(pops, totals):
            f.write("{},{},{}\n".format(pop, total, total / grand))
    with open(hpf_path, "w", newline="") as f:
        f.write("hap,pop,freq\r\n")            # the reference writes through csv.writer: CRLF rows
        for (pop, haplotype), freq in freqs.items():
            f.write("%s,%s,%s\r\n" % (haplotype, pop, freq))


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("-c", "--config", required=False, default="../../conf/minimal-configuration.json",
                        help="Configuration JSON file", type=str)
    produce_hpf(parser.parse_args().config)
