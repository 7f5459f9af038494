"""Deterministic synthetic frequency tables and subject files (shared by the golden generator,
the parity tests and bench.py).  Shapes follow SURVEY.md section 8(d): C2 (single population,
fully typed, unambiguous), C3 (multi-population with race fields), C4 (ambiguous / missing
loci / unknown alleles)."""
import numpy as np

LOCI5 = ["A", "B", "C", "DQB1", "DRB1"]


# --------------------------------------------------------------------------- tables
def parse_hpf(text):
    """-> list of (hap, pop, freq_str) without the header."""
    rows = []
    for line in text.splitlines():
        if not line or line.startswith("hap,"):
            continue
        hap, pop, freq = line.split(",")
        rows.append((hap, pop, freq))
    return rows


def multipop_hpf(base_text, pops, seed, zero_frac=0.4):
    """Perturbed copy of a single-population hpf: freq * LogNormal(0,1), zero_frac of the
    (hap, pop) entries dropped, renormalised per population.  Returns (hpf_text, counts_text)."""
    rng = np.random.RandomState(seed)
    base = parse_hpf(base_text)
    haps = [h for h, _p, _f in base]
    f0 = np.array([float(f) for _h, _p, f in base])
    lines = ["hap,pop,freq\n"]
    for p in pops:
        f = f0 * rng.lognormal(0.0, 1.0, size=len(f0))
        f[rng.rand(len(f0)) < zero_frac] = 0.0
        f = f / f.sum()
        for h, x in zip(haps, f):
            if x > 0:
                lines.append("%s,%s,%s\n" % (h, p, repr(float(x))))
    counts = 1000.0 / np.arange(1, len(pops) + 1) ** 1.1
    tot = counts.sum()
    ctext = "".join("%s,%s,%s\n" % (p, repr(float(c)), repr(float(c / tot))) for p, c in zip(pops, counts))
    return "".join(lines), ctext


def zipf_table(n_full, n_alleles, seed, loci=LOCI5, pops=("CAU",)):
    """Synthetic table: each haplotype is an independent Zipf(s=1.1) draw per locus, frequency
    proportional to 1/rank, normalised.  Returns hpf text."""
    rng = np.random.RandomState(seed)
    cols = []
    for loc, na in zip(loci, n_alleles):
        w = 1.0 / np.arange(1, na + 1) ** 1.1
        w /= w.sum()
        cols.append(rng.choice(na, size=int(n_full * 1.3), p=w))
    tup = np.stack(cols, axis=1)
    _u, first = np.unique(tup, axis=0, return_index=True)
    tup = tup[np.sort(first)][:n_full]
    n = len(tup)
    lines = ["hap,pop,freq\n"]
    for p_i, p in enumerate(pops):
        f = 1.0 / np.arange(1, n + 1)
        if p_i:
            f = f * rng.lognormal(0.0, 1.0, size=n)
        f /= f.sum()
        names = ["~".join("%s*%02d:%02d" % (loc, a // 60 + 1, a % 60 + 1) for loc, a in zip(loci, row)) for row in tup]
        lines.extend("%s,%s,%s\n" % (h, p, repr(float(x))) for h, x in zip(names, f))
    return "".join(lines)


class Table:
    """Sampling helper over one population column of an hpf text."""

    def __init__(self, hpf_text, pop=None, loci=LOCI5):
        rows = parse_hpf(hpf_text)
        if pop is None:
            pop = rows[0][1]
        self.loci = loci
        self.haps = []
        f = []
        for h, p, x in rows:
            if p == pop:
                al = {a.split("*")[0]: a for a in h.split("~")}
                self.haps.append([al[l] for l in loci])
                f.append(float(x))
        f = np.array(f)
        self.p = f / f.sum()
        self.alleles = [sorted({h[i] for h in self.haps}) for i in range(len(loci))]


# --------------------------------------------------------------------------- subjects
def _gl(loci_sides):
    return "^".join("/".join(a) + "+" + "/".join(b) for a, b in loci_sides)


def typed_subjects(table, n, seed, races=None, prefix="S"):
    """Fully typed, unambiguous; the two haplotypes are drawn proportional to frequency so
    Plan A hits (C2).  races: None, or a list of race-field generators cycled per subject."""
    rng = np.random.RandomState(seed)
    idx = rng.choice(len(table.haps), size=(n, 2), p=table.p)
    flip = rng.rand(n, len(table.loci)) < 0.5
    out = []
    for s in range(n):
        h1, h2 = table.haps[idx[s, 0]], table.haps[idx[s, 1]]
        sides = []
        for l in range(len(table.loci)):
            a, b = (h2[l], h1[l]) if flip[s, l] else (h1[l], h2[l])
            sides.append(([a], [b]))
        line = "%s%d,%s" % (prefix, s, _gl(sides))
        if races is not None:
            line += "," + races[s % len(races)]
        out.append(line + "\n")
    return out


def messy_subjects(table, n, seed, max_amb=4, p_missing=0.25, p_unknown=0.05, p_random=0.15,
                   races=None, prefix="M"):
    """Ambiguous / missing-loci / unknown-allele subjects (C4-like, small enough for the CPU
    oracle): each locus side lists the true allele plus up to max_amb-1 others; loci are
    dropped with probability p_missing (at least one kept); alleles are replaced by a name
    absent from the table with probability p_unknown; with probability p_random a side is a
    random allele instead of one from a table haplotype (pushes subjects into Plan B/C)."""
    rng = np.random.RandomState(seed)
    out = []
    nl = len(table.loci)
    for s in range(n):
        i1, i2 = rng.choice(len(table.haps), size=2, p=table.p)
        h1, h2 = list(table.haps[i1]), list(table.haps[i2])
        keep = rng.rand(nl) >= p_missing
        if not keep.any():
            keep[rng.randint(nl)] = True
        sides = []
        for l in range(nl):
            if not keep[l]:
                continue
            pair = []
            for h in (h1, h2):
                a = h[l]
                if rng.rand() < p_random:
                    a = table.alleles[l][rng.randint(len(table.alleles[l]))]
                if rng.rand() < p_unknown:
                    a = "%s*99:%02d" % (table.loci[l], rng.randint(1, 4))
                lst = [a]
                for _ in range(rng.randint(0, max_amb)):
                    b = table.alleles[l][rng.randint(len(table.alleles[l]))]
                    if b not in lst:
                        lst.append(b)
                rng.shuffle(lst)
                pair.append(lst)
            if rng.rand() < 0.5:
                pair.reverse()
            sides.append((pair[0], pair[1]))
        line = "%s%d,%s" % (prefix, s, _gl(sides))
        if races is not None:
            line += "," + races[rng.randint(len(races))]
        out.append(line + "\n")
    return out


def heavy_subjects(table, n, seed, sizes=(6, 12, 20, 40), p_missing=(0.4, 0.3, 0.2, 0.1), p_unknown=0.05,
                   races=None, prefix="H"):
    """Highly ambiguous subjects: every locus side lists a_l alleles drawn from
    `sizes` -- the true allele, then alleles sharing its 2-digit family (serology / MAC-like
    groups), then random ones -- so the product of the list sizes straddles the 100,000-option
    threshold; 0-3 loci are missing (probabilities p_missing) and p_unknown of the listed
    alleles are names absent from the table."""
    rng = np.random.RandomState(seed)
    nl = len(table.loci)
    fam = []
    for l in range(nl):
        d = {}
        for a in table.alleles[l]:
            d.setdefault(a.split(":")[0], []).append(a)
        fam.append(d)
    out = []
    for s in range(n):
        i1, i2 = rng.choice(len(table.haps), size=2, p=table.p)
        hh = (table.haps[i1], table.haps[i2])
        n_miss = int(rng.choice(len(p_missing), p=p_missing))
        drop = set(rng.choice(nl, size=min(n_miss, nl - 1), replace=False).tolist()) if n_miss else set()
        sides = []
        for l in range(nl):
            if l in drop:
                continue
            pair = []
            for h in hh:
                want = int(sizes[rng.randint(len(sizes))])
                lst = [h[l]]
                pool = [a for a in fam[l][h[l].split(":")[0]] if a != h[l]]
                rng.shuffle(pool)
                lst.extend(pool[: want - 1])
                while len(lst) < min(want, len(table.alleles[l])):
                    b = table.alleles[l][rng.randint(len(table.alleles[l]))]
                    if b not in lst:
                        lst.append(b)
                lst = [("%s*99:%02d" % (table.loci[l], rng.randint(1, 9)) if rng.rand() < p_unknown else a) for a in lst]
                lst = list(dict.fromkeys(lst))
                rng.shuffle(lst)
                pair.append(lst)
            sides.append((pair[0], pair[1]))
        line = "%s%d,%s" % (prefix, s, _gl(sides))
        if races is not None:
            line += "," + races[rng.randint(len(races))]
        out.append(line + "\n")
    return out


def race_fields(pops):
    """A small cycle of race1,race2 field shapes: both known, one ';' list,
    unknown code, empty."""
    p = list(pops)
    k = len(p)
    return [
        "%s,%s" % (p[0], p[1 % k]),
        "%s;%s,%s" % (p[0], p[2 % k], p[1 % k]),
        "%s,%s" % (p[1 % k], p[1 % k]),
        "XXX,%s" % p[2 % k],
        ",",
        "%s,XXX;YYY" % p[0],
    ]


# --------------------------------------------------------------------------- BASELINE-size tables (arrays)
def multipop_freqs(base_f, n_pops, seed, zero_frac=0.4):
    """SURVEY 8(d) C3 frequencies for an array table: per population base_f * LogNormal(0, 1) with
    zero_frac of the (haplotype, population) entries zeroed, renormalised per population.
    base_f: [N] -> [N][n_pops]."""
    rng = np.random.RandomState(seed)
    n = len(base_f)
    ff = np.empty((n, n_pops), np.float64)
    for j in range(n_pops):
        f = base_f * rng.lognormal(0.0, 1.0, size=n)
        f[rng.rand(n) < zero_frac] = 0.0
        ff[:, j] = f / f.sum()
    return ff


def pop_counts(pops):
    """-> (text of a pops_count_file with counts proportional to a Zipf law, the ratio column as floats)."""
    counts = 1000.0 / np.arange(1, len(pops) + 1) ** 1.1
    tot = counts.sum()
    text = "".join("%s,%s,%s\n" % (p, repr(float(c)), repr(float(c / tot))) for p, c in zip(pops, counts))
    return text, np.array([float(repr(float(c / tot))) for c in counts])


def array_subject_lines(names, full_alleles, p, n, seed, races=None, prefix="S", variants=True):
    """Lines for a table given as arrays (names[l][id-1], full_alleles [N][L]).  Two haplotypes drawn with
    probabilities p; fully typed, one allele per side.  With `variants`, one subject in eight is changed
    in one of these ways (cycled): fully homozygous, homozygous at two loci, one allele replaced by a name
    absent from the table, one allele replaced by a random table allele (recombinant: usually Plan B),
    both of the last two; so the batch also exercises the hand-over from the warp-per-subject kernels."""
    rng = np.random.RandomState(seed)
    L = full_alleles.shape[1]
    idx = rng.choice(len(p), size=(n, 2), p=p)
    flip = rng.rand(n, L) < 0.5
    h1, h2 = full_alleles[idx[:, 0]], full_alleles[idx[:, 1]]
    a = np.where(flip, h2, h1)
    b = np.where(flip, h1, h2)
    loci = [nm[0].split("*")[0] for nm in names]
    out = []
    for s in range(n):
        sa = [names[l][a[s, l] - 1] for l in range(L)]
        sb = [names[l][b[s, l] - 1] for l in range(L)]
        if variants and s % 8 == 7:
            v = (s // 8) % 5
            if v == 0:
                sb = list(sa)
            elif v == 1:
                for l in rng.choice(L, size=2, replace=False):
                    sb[l] = sa[l]
            if v in (2, 4):
                l = rng.randint(L)
                sa[l] = "%s*99:%02d" % (loci[l], rng.randint(1, 4))
            if v in (3, 4):
                l = rng.randint(L)
                sb[l] = names[l][rng.randint(len(names[l]))]
        line = "%s%d,%s" % (prefix, s, "^".join(x + "+" + y for x, y in zip(sa, sb)))
        if races is not None:
            line += "," + races[s % len(races)]
        out.append(line + "\n")
    return out


class _LazyHaps(object):
    def __init__(self, names, fa):
        self.names, self.fa = names, fa

    def __len__(self):
        return len(self.fa)

    def __getitem__(self, i):
        row = self.fa[i]
        return [self.names[l][int(row[l]) - 1] for l in range(len(self.names))]


class ArrayTable(object):
    """The sampling interface of `Table` over an array table (no per-haplotype Python lists)."""

    def __init__(self, names, full_alleles, p):
        self.loci = [nm[0].split("*")[0] for nm in names]
        self.haps = _LazyHaps(names, full_alleles)
        self.p = p
        self.alleles = [list(n) for n in names]


def zipf_arrays(n_full, n_alleles, seed, loci):
    """zipf_table as arrays: -> (names per locus, full_alleles uint16 [N][L] 1-based, base freq [N])."""
    rng = np.random.RandomState(seed)
    cols = []
    for na in n_alleles:
        w = 1.0 / np.arange(1, na + 1) ** 1.1
        w /= w.sum()
        cols.append(rng.choice(na, size=int(n_full * 1.3) + 100, p=w))
    tup = np.stack(cols, axis=1)
    _u, first = np.unique(tup, axis=0, return_index=True)
    tup = tup[np.sort(first)][:n_full]
    n = len(tup)
    f = 1.0 / np.arange(1, n + 1)
    f /= f.sum()
    names = [["%s*%02d:%02d" % (loc, a // 60 + 1, a % 60 + 1) for a in range(na)] for loc, na in zip(loci, n_alleles)]
    # only alleles that occur are table alleles: re-index so that ids are dense and names stay sorted
    fa = np.zeros((n, len(loci)), np.uint16)
    out_names = []
    for l in range(len(loci)):
        used = np.unique(tup[:, l])
        remap = np.zeros(n_alleles[l], np.int64)
        remap[used] = np.arange(1, len(used) + 1)
        fa[:, l] = remap[tup[:, l]]
        out_names.append([names[l][a] for a in used])
    return out_names, fa, f


def wide_subjects(table, n, seed, sizes=(7, 8, 9), races=None, prefix="W"):
    """Fully typed subjects whose every locus side lists `sizes` alleles (the true one first, then random table
    alleles): Cartesian products of 16,807-59,049 candidates per phase and side, all below the 100,000-option
    threshold -- the shape for which the reference spends seconds per subject and for which
    the cooperative slot pass spreads one subject over up to 2^L CTAs."""
    rng = np.random.RandomState(seed)
    nl = len(table.loci)
    out = []
    for s in range(n):
        i1, i2 = rng.choice(len(table.haps), size=2, p=table.p)
        hh = (table.haps[i1], table.haps[i2])
        sides = []
        for l in range(nl):
            pair = []
            for h in hh:
                want = int(sizes[rng.randint(len(sizes))])
                lst = [h[l]]
                while len(lst) < min(want, len(table.alleles[l])):
                    b = table.alleles[l][rng.randint(len(table.alleles[l]))]
                    if b not in lst:
                        lst.append(b)
                rng.shuffle(lst)
                pair.append(lst)
            sides.append((pair[0], pair[1]))
        line = "%s%d,%s" % (prefix, s, _gl(sides))
        if races is not None:
            line += "," + races[rng.randint(len(races))]
        out.append(line + "\n")
    return out


def subset_subjects(table, n, seed, subsets, races=None, prefix="K", amb=0):
    """Subjects typed at the loci of one of `subsets` (lists of 0-based locus indices, cycled), both haplotypes
    drawn from the table so Plan A hits; amb > 0 adds up to `amb` extra table alleles per side (Plan_A_Matrix
    cases: the typed-locus pattern decides whether the reference imputes the subject at all)."""
    rng = np.random.RandomState(seed)
    idx = rng.choice(len(table.haps), size=(n, 2), p=table.p)
    out = []
    for s in range(n):
        h1, h2 = table.haps[idx[s, 0]], table.haps[idx[s, 1]]
        sides = []
        for l in subsets[s % len(subsets)]:
            a, b = [h1[l]], [h2[l]]
            if rng.rand() < 0.5:
                a, b = b, a
            for side in (a, b):
                for _ in range(rng.randint(0, amb + 1) if amb else 0):
                    x = table.alleles[l][rng.randint(len(table.alleles[l]))]
                    if x not in side:
                        side.append(x)
            sides.append((a, b))
        line = "%s%d,%s" % (prefix, s, _gl(sides))
        if races is not None:
            line += "," + races[s % len(races)]
        out.append(line + "\n")
    return out
