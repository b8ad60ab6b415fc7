"""Replicate construction in array form: resampling of alignment records and the per-replicate multinomial bin
statistics, bit-exact against the reference.

The reference builds, per bootstrap replicate, a fresh graph of Python objects
(``Traversome.sample_sub_paths`` traversome.py:1636-1699 -> ``generate_multinomial_bin_stats`` :1305-1384 with
``_identify_bins`` :1386-1462, ``_generate_align_len_id_lookup_table`` :1809-1836,
``get_averaged_possible_alignment_start_points`` :1492-1521 and ``get_num_of_possible_alignment_start_points``
:1523-1579) and hands it to ``PathMultinomialModel``.  ``RecordPool`` holds the same information as flat arrays and
produces, for a sample of records, the reference's bins -- the same ranges, id ranges, ``num_matched`` and
``num_possible_X`` (integer sum / float count: identical bits) -- and from them the lane (weight column, C, N) that the
CUDA solver consumes.  The random draws replay ``random.choices`` / ``random.sample`` of the reference, so a run with the
same seed sees the same replicates.
"""
import math

import numpy as np


class RecordPool(object):
    """rows = read paths in ``all_sub_paths`` order.

    record_ids[i]   sorted ids of the valid records (``records_pool_sorted``)
    row_of[i]       read path of record i (``records_pool_to_sbp_ids``)
    align_len[i]    ``align_len_at_path_map[record_ids[i]]``
    multi[r]        read path r has more than one contig
    inner_len[r], full_len[r]                       ``SubPathInfo.inner_len / full_len``
    left_room[r] = left contig length - overlap, right_room[r] likewise (multi-contig paths; :1558-1573)
    """

    def __init__(self, record_ids, row_of, align_len, multi, inner_len, full_len, left_room, right_room):
        self.record_ids = np.asarray(record_ids, dtype=np.int64)
        self.row_of = np.asarray(row_of, dtype=np.int64)
        self.align_len = np.asarray(align_len, dtype=np.int64)
        self.multi = np.asarray(multi, dtype=bool)
        self.inner_len = np.asarray(inner_len, dtype=np.int64)
        self.full_len = np.asarray(full_len, dtype=np.int64)
        self.left_room = np.asarray(left_room, dtype=np.int64)
        self.right_room = np.asarray(right_room, dtype=np.int64)
        self.n_records = int(self.record_ids.size)
        self.n_rows = int(self.multi.size)
        if np.any(np.diff(self.record_ids) <= 0):
            raise ValueError("record_ids must be strictly increasing (records_pool_sorted)")
        self._index_list = list(range(self.n_records))
        # _identify_bins: min / max alignment length a read path can produce (:1391-1392)
        self.min_len = np.where(self.multi, self.inner_len + 2, 1)
        self.max_len = self.full_len

    # ------------------------------------------------------------------ construction from reference objects
    @classmethod
    def from_traversome(cls, trv):
        """From a live ``Traversome`` object after ``_prepare_for_sampling`` (traversome.py:1622-1634)."""
        if not getattr(trv, "records_pool_sorted", None):
            trv._prepare_for_sampling()
        paths = list(trv.all_sub_paths.items())
        multi, inner, full, lroom, rroom = [], [], [], [], []
        for path, spi in paths:
            multi.append(len(path) > 1)
            inner.append(spi.inner_len)
            full.append(spi.full_len)
            if len(path) > 1:                       # :1558-1567
                (ln1, le1), (ln2, le2) = path[0], path[1]
                (rn1, re1), (rn2, re2) = path[-1], path[-2]
                li, ri = trv.graph.vertex_info[ln1], trv.graph.vertex_info[rn1]
                lroom.append(li.len - li.connections[le1][(ln2, not le2)])
                rroom.append(ri.len - ri.connections[not re1][(rn2, re2)])
            else:
                lroom.append(0)
                rroom.append(0)
        ids = list(trv.records_pool_sorted)
        return cls(ids, [trv.records_pool_to_sbp_ids[r] for r in ids], [trv.align_len_at_path_map[r] for r in ids],
                   multi, inner, full, lroom, rroom)

    @classmethod
    def from_tables(cls, fixture):
        """From plain tables (JSON-able): fixture["read_paths"] = per read path {path, inner_len, full_len, left_len,
        left_overlap, right_len, right_overlap}, fixture["records"] = {pool_sorted, row_of_record, align_len} -- the
        array form of what `from_traversome` reads off the reference's objects."""
        rp = fixture["read_paths"]
        multi = [len(r["path"]) > 1 for r in rp]
        lroom = [r.get("left_len", 0) - r.get("left_overlap", 0) for r in rp]
        rroom = [r.get("right_len", 0) - r.get("right_overlap", 0) for r in rp]
        rec = fixture["records"]
        return cls(rec["pool_sorted"], rec["row_of_record"], rec["align_len"], multi,
                   [r["inner_len"] for r in rp], [r["full_len"] for r in rp], lroom, rroom)

    # ------------------------------------------------------------------ sampling (replays the reference's draws)
    def sample_bootstrap(self, rng, size=None):
        """Indices into the pool; the same draws as ``random.choices(records_pool_sorted, k=size)`` (:1652)."""
        size = self.n_records if size is None else int(size)
        return np.asarray(rng.choices(self._index_list, k=size), dtype=np.int64)

    def sample_jackknife(self, rng, jackknife_size):
        """``random.sample(range(N), k=N - jackknife_size)`` (:1655)."""
        return np.asarray(rng.sample(range(self.n_records), k=self.n_records - int(jackknife_size)), dtype=np.int64)

    # ------------------------------------------------------------------ bins of one replicate
    def _starts(self, lens, row):
        """get_num_of_possible_alignment_start_points for every length in `lens` (:1523-1579), exact integers."""
        if self.multi[row]:
            m = lens - (self.inner_len[row] + 2) + 1
            return m - np.maximum(m - self.left_room[row], 0) - np.maximum(m - self.right_room[row], 0)
        return self.full_len[row] - lens + 1

    def _ranges(self, rows_present, lens_sorted):
        """_identify_bins (:1386-1462) for the sampled read paths; returns (ranges, legal_rows)."""
        points = {}
        legal = []
        for r in rows_present:
            lo, hi = int(self.min_len[r]), int(self.max_len[r])
            if lo > hi:
                continue                                            # "deleting illegal path"
            legal.append(r)
            if lo == hi:
                points.setdefault(lo, []).append(1)
            else:
                points.setdefault(lo, []).append(0)
                points.setdefault(hi, []).append(2)
        ranges = []
        last_type = False
        sorted_points = sorted(points)
        for go_p, point in enumerate(sorted_points):
            types = points[point]
            is_start = any(t < 2 for t in types)
            is_end = any(t > 0 for t in types)
            if is_start:
                if ranges and last_type:
                    ranges[-1][1] = point - 1
                ranges.append([point, None])
                last_type = True
            if is_end:
                ranges[-1][-1] = point
                last_type = False
                nxt = go_p + 1
                if nxt < len(sorted_points):
                    if not (sorted_points[nxt] == point + 1 and is_start):
                        ranges.append([point + 1, None])
                        last_type = True
        kept = []
        i, j, n = 0, 0, len(lens_sorted)
        while i < len(ranges) and j < n:
            if ranges[i][1] >= lens_sorted[j] >= ranges[i][0]:
                kept.append(ranges[i])
                i += 1
            elif lens_sorted[j] > ranges[i][1]:
                i += 1
            else:
                j += 1
        return kept, legal

    def bins_of(self, sample, masked_rows=()):
        """The reference's bins_list for the records `sample` (indices into the pool, duplicates allowed).

        Returns (bins, observed_rows): bins = list of dicts(min_len, max_len, min_id, max_id, rows int64[],
        num_matched int64[], X float64[]) with the read paths of each range in the reference's (first appearance) order;
        observed_rows = the read paths of the replicate's ``all_sub_paths`` (sampled, not masked, legal)."""
        sample = np.asarray(sample, dtype=np.int64)
        rows_all = self.row_of[sample]
        if len(masked_rows):
            keep = ~np.isin(rows_all, np.asarray(list(masked_rows), dtype=np.int64))
            sample, rows_all = sample[keep], rows_all[keep]
        if sample.size == 0:
            return [], np.zeros(0, dtype=np.int64)
        # sorted by (length, record id) like :1696-1698; duplicates keep their multiplicity
        order = np.lexsort((self.record_ids[sample], self.align_len[sample]))
        lens = self.align_len[sample][order]
        rows = rows_all[order]
        present = np.unique(rows_all)                                # all_sub_paths order = increasing row
        ranges, legal = self._ranges(present.tolist(), lens)
        legal_mask = np.zeros(self.n_rows, dtype=bool)
        legal_mask[legal] = True
        lo_len, hi_len, max_id = int(lens[0]), int(lens[-1]), int(lens.size - 1)
        out = []
        for min_len, max_len in ranges:
            # _get_id_range_in_increasing_lengths over _generate_align_len_id_lookup_table (:1809-1842)
            left_id = int(np.searchsorted(lens, min_len, side="left")) if lo_len <= min_len <= hi_len else 0
            right_id = int(np.searchsorted(lens, max_len, side="right")) - 1 if lo_len <= max_len <= hi_len else max_id
            if left_id > right_id:
                continue                                            # "Remove range"
            seg_rows = rows[left_id:right_id + 1]
            seg_lens = lens[left_id:right_id + 1]
            ok = legal_mask[seg_rows]
            uniq, first, counts = np.unique(seg_rows[ok], return_index=True, return_counts=True)
            by_first = np.argsort(first, kind="stable")
            uniq, counts = uniq[by_first], counts[by_first]
            n_rec = right_id - left_id + 1
            X = np.array([int(self._starts(seg_lens, int(r)).sum()) / float(n_rec) for r in uniq], dtype=np.float64)
            out.append({"min_len": int(min_len), "max_len": int(max_len), "min_id": left_id, "max_id": right_id,
                        "rows": uniq.astype(np.int64), "num_matched": counts.astype(np.int64), "X": X})
        return out, np.asarray(legal, dtype=np.int64)

    # ------------------------------------------------------------------ the same on the device, a batch of replicates per call
    def device_pool(self, device=0):
        """The pool on the device (include/tvs_pool.h, csrc/tvs_pool.cu); created once, reused by `lanes_on_device`."""
        from . import _cabi
        pool = getattr(self, "_device_pool", None)
        if pool is None or pool[0] != device:
            if pool is not None:
                pool[1].close()
            pool = self._device_pool = (device, _cabi.DevicePool(self.record_ids, self.row_of, self.align_len, self.multi, self.inner_len,
                                                                 self.full_len, self.left_room, self.right_room, device=device))
        return pool[1]

    def lanes_on_device(self, samples, masked_rows=(), device=0):
        """`lane_of` for a BATCH of replicates in one device call: samples = list of index arrays (one per replicate, the
        draws of sample_bootstrap / sample_jackknife).  Returns (w int32 [n_rows, B], C [B], N [B], observed bool [n_rows, B])
        -- w, N and observed identical to `lane_of`, C equal to rounding (its terms are summed in another order)."""
        n_draw = max(len(s) for s in samples)
        block = np.full((len(samples), n_draw), -1, dtype=np.int32)
        for b, s_ in enumerate(samples):
            block[b, :len(s_)] = s_
        masked = None
        if len(masked_rows):
            masked = np.zeros(self.n_rows, dtype=np.uint8)
            masked[np.asarray(list(masked_rows), dtype=np.int64)] = 1
        return self.device_pool(device).lanes(block, masked)

    def lane_of(self, sample, masked_rows=()):
        """(w int32[n_rows], C, N, observed bool[n_rows]) of one replicate: the pruning rule of
        ModelGenerator.py:138-151 (bins with num_possible_X < 1 dropped) applied to bins_of(sample)."""
        w = np.zeros(self.n_rows, dtype=np.int32)
        observed = np.zeros(self.n_rows, dtype=bool)
        C, N = 0.0, 0
        bins, obs_rows = self.bins_of(sample, masked_rows)
        observed[obs_rows] = True
        for b in bins:
            for r, n, X in zip(b["rows"], b["num_matched"], b["X"]):
                if X < 1:
                    continue
                w[r] += n
                C += math.log(X) * int(n)
                N += int(n)
        return w, C, N, observed
