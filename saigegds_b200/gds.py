"""Host half of the genotype ingestion without SeqArray / gdsfmt.

The reference opens the GDS file with SeqArray, filters samples and variants (R/saige_main.r:305-333) and pulls the
genotypes through `SeqArray:::.seqGet2bGeno` or `seqApply` (:395-421).  Here the file is read directly: a SeqArray GDS file
keeps every array node as a sequence of independent xz streams (`LZMA_RA` random-access blocks); the streams are found by
their magic number, decompressed in parallel, and the nodes this path needs are recognised by content:

  variant.id     int32 1..n_variant                      -> n_variant
  sample.id      n_sample length-prefixed strings, or int32 ids
  genotype/data  bit2 allele indices [variant][sample][ploidy = 2]: the first run of consecutive streams (blocks of >= 64 KB)
                 whose sizes add up to n_variant * n_sample / 2 bytes

The decompressed `genotype/data` bytes are exactly what `sgb_store_gds_geno` (csrc/store.cu: gds_to_dosage_kernel) takes;
2-bit packing, allele counts and the MAF / missing-rate filter then run on the GPU.  Neither gdsfmt's block directory nor its
node tree is parsed (their sources are not part of the reference tree): files whose genotype node is not LZMA_RA-compressed
bit2 data are rejected with an error instead of being guessed at.
"""
from __future__ import annotations

import lzma
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

XZ_MAGIC = b"\xfd7zXZ\x00"
GDS_MAGIC = b"COREARRAYx0A"


class GdsFormatError(ValueError):
    pass


def _try_stream(buf: bytes, pos: int):
    d = lzma.LZMADecompressor(format=lzma.FORMAT_XZ)
    try:
        out = d.decompress(buf[pos:])
    except lzma.LZMAError:
        return None
    if not d.eof:
        return None
    return pos, len(buf) - pos - len(d.unused_data), out


def xz_streams(buf: bytes, threads: int = 8):
    """[(offset, compressed length, decompressed bytes)] of every complete xz stream in `buf`, in file order.  Candidate
    offsets inside an earlier stream's compressed range are skipped (a magic number can occur in compressed data)."""
    cand = []
    pos = buf.find(XZ_MAGIC)
    while pos >= 0:
        cand.append(pos)
        pos = buf.find(XZ_MAGIC, pos + 1)
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:          # lzma releases the GIL
        res = list(ex.map(lambda p: _try_stream(buf, p), cand))
    out, end = [], 0
    for r in res:
        if r is None or r[0] < end:
            continue
        out.append(r)
        end = r[0] + r[1]
    return out


def _parse_strings(raw: bytes, limit: int = 1 << 26):
    """Length-prefixed strings (LEB128 length, then the bytes); None if `raw` is not such a list."""
    out, i, n = [], 0, len(raw)
    while i < n:
        ln, shift = 0, 0
        while True:
            if i >= n:
                return None
            b = raw[i]
            i += 1
            ln |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        if i + ln > n or len(out) > limit:
            return None
        out.append(raw[i:i + ln])
        i += ln
    return out


@dataclass
class GdsGenotypes:
    path: str
    n_sample: int
    n_variant: int
    sample_id: np.ndarray          # strings or int32
    variant_id: np.ndarray         # int32, 1-based
    allele_bits: np.ndarray        # uint8, the decompressed genotype/data node
    stream_offsets: list           # file offsets of the genotype streams (diagnostics)


def read_gds_genotypes(path: str, n_sample: int | None = None, n_variant: int | None = None, threads: int = 8) -> GdsGenotypes:
    buf = open(path, "rb").read()
    if not buf.startswith(GDS_MAGIC):
        raise GdsFormatError("%s is not a GDS (CoreArray) file" % path)
    streams = xz_streams(buf, threads)
    if not streams:
        raise GdsFormatError("no LZMA_RA (xz) streams found: the genotype node must be compressed with LZMA_RA")
    # variant.id: the first int32 stream counting 1..n
    variant_id = None
    for _, _, raw in streams:
        if len(raw) % 4 == 0 and len(raw) >= 4:
            a = np.frombuffer(raw, dtype="<i4")
            if a[0] == 1 and (n_variant is None or len(a) == n_variant) and np.array_equal(a, np.arange(1, len(a) + 1, dtype=np.int32)):
                variant_id = a
                break
    if variant_id is None and n_variant is None:
        raise GdsFormatError("variant.id (int32 1..n) not found; pass n_variant")
    n_variant = int(n_variant if n_variant is not None else len(variant_id))
    if variant_id is None:
        variant_id = np.arange(1, n_variant + 1, dtype=np.int32)
    # sample.id: the first stream that is a list of strings (or, failing that, use n_sample as given)
    sample_id = None
    for _, _, raw in streams:
        if len(raw) >= 65536 and len(raw) % 4 == 0:
            continue
        s = _parse_strings(raw)
        if s is not None and len(s) > 0 and (n_sample is None or len(s) == n_sample) and all(len(x) > 0 for x in s[:16]):
            try:
                sample_id = np.array([x.decode("utf-8") for x in s])
            except UnicodeDecodeError:
                continue
            break
    if n_sample is None:
        if sample_id is None:
            raise GdsFormatError("sample.id not recognised; pass n_sample")
        n_sample = len(sample_id)
    n_sample = int(n_sample)
    if sample_id is None:
        sample_id = np.arange(1, n_sample + 1, dtype=np.int32)
    want = (n_variant * n_sample * 2 * 2 + 7) // 8                         # bit2 x ploidy 2
    sizes = [len(s[2]) for s in streams]
    for i in range(len(streams)):
        if sizes[i] < min(65536, want):
            continue
        tot, j = 0, i
        while j < len(streams) and tot < want:
            tot += sizes[j]
            j += 1
        if tot == want:
            bits = np.frombuffer(b"".join(s[2] for s in streams[i:j]), dtype=np.uint8)
            return GdsGenotypes(path, n_sample, n_variant, sample_id, variant_id, bits, [s[0] for s in streams[i:j]])
    raise GdsFormatError("genotype/data not found: no run of xz streams holds %d bytes (n_variant %d x n_sample %d, bit2, ploidy 2)"
                         % (want, n_variant, n_sample))


def store_from_gds(ctx, path: str, sample_id=None, maf: float = float("nan"), missing_rate: float = float("nan"), threads: int = 8):
    """seqOpen + seqSetFilter(sample.id) + seqSetFilterCond(maf, missing.rate) + .seqGet2bGeno + saige_store_2b_geno
    (R/saige_main.r:305-333, 395-437): reads the file, selects the samples listed in `sample_id` (file order is kept, as
    SeqArray does), filters the variants on the device and stores the packed genotypes in `ctx`.
    Returns (GdsGenotypes, dict from Context.store_gds_geno with `variant_sel`, allele counts, lut, diag, `sample_sel`)."""
    g = read_gds_genotypes(path, threads=threads)
    sel = None
    if sample_id is not None:
        want = set(np.asarray(sample_id).astype(g.sample_id.dtype).tolist())
        sel = np.flatnonzero(np.array([s in want for s in g.sample_id.tolist()])).astype(np.int32)
        if len(sel) == 0:
            raise ValueError("No sample in the GDS file matches 'sample.id'.")
    r = ctx.store_gds_geno(g.allele_bits, g.n_sample, g.n_variant, sample_sel=sel, maf=maf, missing_rate=missing_rate)
    r["sample_sel"] = sel
    return g, r
