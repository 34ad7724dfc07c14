"""Hand-built cells shared by the emulator tier and the GPU tier (test infrastructure)."""
import numpy as np


def star_cells(rng, n_cells, n_genes):
    """cells made of 1-Hamming 'stars' (a centre UMI + 2..6 of its single-base substitutions) over a
    few overlapping transcript sets: components of 6..30 (class, UMI) vertices for the warp-cooperative cover"""
    cells = []
    for _ in range(n_cells):
        recs = []
        for _star in range(int(rng.integers(3, 7))):
            centre = int(rng.integers(0, 1 << 24))
            g = int(rng.integers(0, n_genes - 2))
            labels = [[3 * g], [3 * g, 3 * g + 1], [3 * g, 3 * g + 1, 3 * g + 4], [3 * g + 1, 3 * g + 4], [3 * g + 4, 3 * g + 5]]
            leaves = rng.choice(36, size=int(rng.integers(2, 7)), replace=False)
            umis = [centre] + [centre ^ ((int(k) % 3 + 1) << (2 * (int(k) // 3))) for k in leaves]
            for u in umis:
                for _r in range(int(rng.integers(1, 4))):
                    recs.append((u, labels[int(rng.integers(0, len(labels)))]))
        for _bg in range(120):   # background so the cell is not tiny
            recs.append((int(rng.integers(0, 1 << 24)), [3 * int(rng.integers(0, n_genes))]))
        rng.shuffle(recs)
        cells.append(recs)
    return cells




def long_label_cells():
    """cells whose eq-class labels hold 32..40 transcripts: (n_genes, tid_to_gid, cells)"""
    rng = np.random.default_rng(3)
    n_genes, per = 20, 50
    t2g = np.repeat(np.arange(n_genes, dtype=np.uint32), per)
    cells = []
    for _ in range(4):
        recs = []
        for _u in range(30):
            u = int(rng.integers(0, 1 << 24))
            g = int(rng.integers(0, n_genes - 1))
            a = list(range(g * per, g * per + 40))
            bsub = list(range(g * per + 3, g * per + 38))
            csub = list(range(g * per + 30, g * per + 62))          # spills into the next gene
            for lab in (a, bsub, csub)[: int(rng.integers(2, 4))]:
                for _r in range(int(rng.integers(1, 4))):
                    recs.append((u, lab))
            if rng.random() < 0.5:
                recs.append((u ^ 1, a))                                # a 1-Hamming neighbour
        for _bg in range(110):
            recs.append((int(rng.integers(0, 1 << 24)), [int(rng.integers(0, n_genes * per))]))
        rng.shuffle(recs)
        cells.append(recs)
    return n_genes, t2g, cells


def random_pug_cells(rng, n_cells, n_tx, umi_bits, n_labels, max_label, recs_lo, recs_hi, dup=3):
    """adversarial small cells for the PUG kernels: a small UMI space (umi_bits) and a small pool of
    overlapping transcript labels make components of every size (2 .. beyond 32), many cover ties,
    directed and bidirected edges (skewed read counts)"""
    cells = []
    for _ in range(n_cells):
        pool = []
        for _l in range(n_labels):
            k = int(rng.integers(1, max_label + 1))
            start = int(rng.integers(0, max(1, n_tx - k)))
            if rng.random() < 0.5:
                lab = list(range(start, start + k))                      # a contiguous run (shared prefixes / suffixes)
            else:
                lab = sorted(set(int(x) for x in rng.integers(0, n_tx, size=k)))
            pool.append(lab)
        recs = []
        n = int(rng.integers(recs_lo, recs_hi))
        while len(recs) < n:
            u = int(rng.integers(0, 1 << umi_bits))
            lab = pool[int(rng.integers(0, len(pool)))]
            reps = 1 + int(rng.geometric(1.0 / dup)) if rng.random() < 0.5 else 1
            recs.extend([(u, lab)] * reps)
        rng.shuffle(recs)
        cells.append(recs[:n])
    return cells


def record_shape_cells(rng):
    """cr-like phase 1 (afq_kernels.cuh), the record shapes its two forms must agree on: cells dense in multi-gene records, a
    record of 300 alignments, records that repeat a gene out of order (unsorted refs), alignment-free records.
    3 transcripts per gene; returns (n_genes, tid_to_gid, cells)."""
    import numpy as np
    n_genes = 400
    t2g = np.repeat(np.arange(n_genes, dtype=np.uint32), 3)
    cells = []
    # 200 records of three different genes each
    c = []
    for i in range(200):
        gs = rng.choice(n_genes, 3, replace=False)
        c.append((int(rng.integers(0, 60)), sorted(int(3 * g + rng.integers(0, 3)) for g in gs)))
    cells.append(c)
    # one very long multi-gene record among single-alignment ones
    c = [(int(rng.integers(0, 50)), [int(rng.integers(0, 3 * n_genes))]) for _ in range(220)]
    c.insert(17, (7, sorted(int(x) for x in rng.choice(3 * n_genes, 300, replace=False))))
    cells.append(c)
    # genes repeated out of order inside a record (unsorted refs), duplicates of whole records, records without alignments
    c = []
    for i in range(300):
        g1, g2 = (int(x) for x in rng.choice(n_genes, 2, replace=False))
        kind = i % 4
        refs = [3 * g1, 3 * g2, 3 * g1 + 1] if kind == 0 else ([3 * g2 + 2, 3 * g1] if kind == 1 else ([] if kind == 2 else [3 * g1 + 2, 3 * g1]))
        c.append((int(rng.integers(0, 40)), refs))
    cells.append(c)
    # a mid-size cell (a larger arena): every fourth record multi-gene
    c = []
    for i in range(3000):
        g1 = int(rng.integers(0, n_genes))
        refs = [3 * g1] if i % 4 else sorted([3 * g1, 3 * ((g1 + 1 + int(rng.integers(0, 5))) % n_genes) + 1])
        c.append((int(rng.integers(0, 700)), refs))
    cells.append(c)
    return n_genes, t2g, cells
