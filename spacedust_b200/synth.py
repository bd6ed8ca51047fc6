"""Seeded synthetic prokaryote proteomes: a pool of ancestral protein families,
each proteome = core families + a sample of accessory families, every copy mutated to 30-90 % identity
with short indels, gene order kept in syntenic blocks that are shuffled (and strand-flipped) per
proteome.  FASTA headers are Prodigal style (`>acc_N # start # end # strand # ID=..`) so that the
reference's createsetdb accepts them unchanged (data/createsetdb.sh:121-135)."""
import os

import numpy as np

LETTERS = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
# BLOSUM62 background of the 20 letters (lib/mmseqs/data/blosum62.out line 2)
BACKGROUND = np.array([0.07422, 0.02469, 0.05363, 0.05431, 0.04742, 0.07415, 0.02621, 0.06792, 0.05815, 0.09891,
                       0.02499, 0.04465, 0.03854, 0.03426, 0.05161, 0.05723, 0.05089, 0.07292, 0.01303, 0.03228])
BACKGROUND = BACKGROUND / BACKGROUND.sum()


def _mutate(rng, anc, identity):
    """Point substitutions to the requested identity + geometric indels at rate ~1/50."""
    L = len(anc)
    seq = anc.copy()
    sub = rng.random(L) >= identity
    n = int(sub.sum())
    if n:
        seq[sub] = rng.choice(20, size=n, p=BACKGROUND).astype(np.uint8)
    n_indel = rng.binomial(L, 1.0 / 50.0)
    if n_indel:
        pos = np.sort(rng.integers(0, L, size=n_indel))
        pieces = []
        last = 0
        for p in pos:
            if p < last:
                continue
            pieces.append(seq[last:p])
            ln = int(rng.geometric(0.5))
            if rng.random() < 0.5:
                pieces.append(rng.choice(20, size=ln, p=BACKGROUND).astype(np.uint8))
                last = p
            else:
                last = min(L, p + ln)
        pieces.append(seq[last:])
        seq = np.concatenate(pieces)
    if len(seq) < 30:
        seq = np.concatenate([seq, rng.choice(20, size=30 - len(seq), p=BACKGROUND).astype(np.uint8)])
    return seq


def generate(n_proteomes, genes_per_proteome, seed=20261017, core_fraction=0.3, pool_factor=5):
    """-> list of proteomes; a proteome is a dict(acc=str, names=[header...], seqs=[bytes...])."""
    rng = np.random.default_rng(seed)
    G = genes_per_proteome
    n_fam = max(G, int(pool_factor * G))
    lens = np.clip(np.exp(rng.normal(np.log(280.0), 0.55, size=n_fam)), 40, 2500).astype(np.int64)
    ancestors = [rng.choice(20, size=int(l), p=BACKGROUND).astype(np.uint8) for l in lens]
    n_core = int(round(core_fraction * G))
    out = []
    for pi in range(n_proteomes):
        prng = np.random.default_rng(seed + 1 + pi)
        acc_fams = prng.choice(np.arange(n_core, n_fam), size=G - n_core, replace=False) if n_fam > n_core else np.array([], dtype=np.int64)
        fams = np.sort(np.concatenate([np.arange(n_core), acc_fams]))
        # syntenic blocks of 3..15 consecutive family members, shuffled, random strand
        blocks = []
        i = 0
        while i < len(fams):
            b = int(prng.integers(3, 16))
            blocks.append(fams[i:i + b])
            i += b
        order = prng.permutation(len(blocks))
        acc = "SYN%05d.1" % pi
        names, seqs = [], []
        pos = 1
        gi = 0
        for bi in order:
            blk = blocks[bi]
            strand = 1 if prng.random() < 0.5 else -1
            if strand < 0:
                blk = blk[::-1]
            for f in blk:
                ident = prng.uniform(0.3, 0.9)
                s = _mutate(prng, ancestors[int(f)], ident)
                gi += 1
                start = pos
                end = pos + 3 * len(s) + 2
                pos = end + int(prng.integers(20, 200))
                names.append("%s_%d # %d # %d # %d # ID=%d_%d;partial=00;start_type=ATG;rbs_motif=None;rbs_spacer=None;gc_cont=0.500"
                             % (acc, gi, start, end, strand, pi + 1, gi))
                seqs.append(LETTERS[s].tobytes() + b"*")
        out.append({"acc": acc, "names": names, "seqs": seqs})
    return out


def write_fasta(proteomes, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for p in proteomes:
        path = os.path.join(out_dir, p["acc"] + ".faa")
        with open(path, "wb") as f:
            for n, s in zip(p["names"], p["seqs"]):
                f.write(b">" + n.encode() + b"\n")
                for i in range(0, len(s), 60):
                    f.write(s[i:i + 60] + b"\n")
        paths.append(path)
    return paths


def pack(proteomes):
    """All proteins of all proteomes in file order -> (uint8 residues, uint64 offsets, set id per protein)."""
    seqs = [s for p in proteomes for s in p["seqs"]]
    sets = np.array([i for i, p in enumerate(proteomes) for _ in p["seqs"]], dtype=np.uint32)
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    return buf, off, sets


def set_metadata(proteomes):
    """Per protein (file order): set index, gene index inside the set, strand (1: start < end), plus the lookup
    names `<id>_<geneIdx>_<start>_<end>` createsetdb writes (data/createsetdb.sh:121-135; start > end on the minus strand)."""
    set_of, pos, strand, names = [], [], [], []
    for si, p in enumerate(proteomes):
        for gi, header in enumerate(p["names"]):
            f = header.split(" # ")
            start, end, st = int(f[1]), int(f[2]), int(f[3])
            set_of.append(si)
            pos.append(gi)
            strand.append(1 if st > 0 else 0)
            names.append("%s_%d_%d_%d" % (f[0], gi, start, end) if st > 0 else "%s_%d_%d_%d" % (f[0], gi, end, start))
    return {"set_of": np.array(set_of, dtype=np.uint32), "pos": np.array(pos, dtype=np.uint32),
            "strand": np.array(strand, dtype=np.uint8),
            "set_size": np.array([len(p["names"]) for p in proteomes], dtype=np.uint32), "names": names}
