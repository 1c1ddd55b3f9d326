"""GMAP4's reduce key is the md5 of an UNSEPARATED string concatenation (GMAP4.scala:32):

    key._1 + key._2 + ref._1 + ref._2 + ref._3 + ref._4.mkString("/")      # groupId, chrom, start, stop, strand, values

so ("chr1", 1234567, ...) and ("chr11", 234567, ...) build the same string, and two phase-1 regions with equal strings are
reduced into ONE output row by the reference (contributors pooled; which coordinates survive depends on Spark's shuffle order).
The device path keys rows by (group, chrom index, start, stop, strand), i.e. it never aliases.  This host post-pass over the
(small) output table reports how many key strings are shared by more than one region, so that a run whose result could differ
from the reference's for this reason says so (expected 0)."""
import numpy as np

_STRAND = "*+-"


def gmap4_key_strings(group_ids, chrom_names, group, chrom, start, stop, strand, acc):
    """The reference's key string of every output row (values = [GDouble(AccIndex)] printed by GDouble.toString)."""
    # AccIndex is an integer held in a GDouble: DecimalFormat("#.########") prints it without a fraction (DataTypes.scala:144-147)
    return ["%d%s%d%d%s%d" % (int(group_ids[g]), chrom_names[c], int(s), int(e), _STRAND[int(t)], int(a))
            for g, c, s, e, t, a in zip(group, chrom, start, stop, strand, acc)]


def count_aliased_groups(group_ids, chrom_names, group, chrom, start, stop, strand, acc):
    """Number of key strings shared by two or more distinct output regions."""
    keys = gmap4_key_strings(group_ids, chrom_names, group, chrom, start, stop, strand, acc)
    if not keys:
        return 0
    u, cnt = np.unique(np.array(keys, dtype=object), return_counts=True)
    return int((cnt > 1).sum())
