"""Result tables of the ML-fit path, in the reference's formats.

Mirrors ``Traversome.output_variant_info`` (traversome/traversome.py:738-781), ``output_result_info`` (:669-722) and
``output_sampling_info`` (:724-736): the same columns, 1-based ``fid`` numbering in order of appearance, ``%.4f``
proportions, ``-`` for variants a solution does not use and ``*`` for solutions that were not re-evaluated.  The
functions return the table text (``write_tables`` puts the files into a directory under the reference's file names), so
that the tables of a run through this package can be diffed against the reference's.
"""
import os
from collections import OrderedDict

import numpy as np

from .utils import Criterion


def path_to_gaf_str(path):
    """(('53', True), ('45', False)) -> '>53<45' (traversome/utils.py:1085-1092)."""
    return "".join("%s%s" % (">" if e else "<", n) for n, e in path)


def cid_to_fid_table(vp_unique_results_sorted, variant_proportions):
    """Final (1-based) ids of the candidate variants that appear in any solution (traversome.py:739-752)."""
    cid_to_fid = OrderedDict()
    for tuple_v_prop in vp_unique_results_sorted:
        for cid in tuple_v_prop:
            if cid not in cid_to_fid:
                cid_to_fid[cid] = len(cid_to_fid) + 1
    for cid in variant_proportions:
        if cid not in cid_to_fid:
            cid_to_fid[cid] = len(cid_to_fid) + 1
    return cid_to_fid


def variants_info_tab(cid_to_fid, be_unidentifiable_to, variant_paths, variant_sizes):
    """variants.info.tab (traversome.py:754-761)."""
    lines = ["FID\tCID\tUnidentifiable_to_cid\tCONTIGS\tBASES\tPATH\n"]
    for cid, fid in cid_to_fid.items():
        uid = be_unidentifiable_to.get(cid, cid)
        uid = "-" if uid == cid else uid
        path = variant_paths[cid]
        lines.append("%s\t%s\t%s\t%d\t%s\t%s\n" % (fid, cid, uid, len(path), variant_sizes[cid], path_to_gaf_str(path)))
    return "".join(lines)


def readpath_information_tab(cid_to_fid, read_paths, variant_paths, variant_subpath_counters):
    """readpath.information.tab (traversome.py:763-773).  `read_paths`: ordered mapping read path -> record ids."""
    lines = ["rp_id\t" + "\t".join("FID_%s" % fid for fid in cid_to_fid.values()) + "\tpath\tnum_reads\n"]
    counters = [variant_subpath_counters[variant_paths[cid]] for cid in cid_to_fid]
    for go_rp, (path, record_ids) in enumerate(read_paths.items()):
        cells = "".join("%s\t" % c.get(path, 0) for c in counters)
        lines.append("%d\t%s%s\t%d\n" % (go_rp, cells, path_to_gaf_str(path), len(record_ids)))
    return "".join(lines)


def readpath_record_ids_tab(read_paths):
    """readpath.record_ids.tab (traversome.py:775-778)."""
    lines = ["rp_id\trecord_id\n"]
    for go_rp, record_ids in enumerate(read_paths.values()):
        lines.append("%d\t%s\n" % (go_rp, ",".join(str(r) for r in record_ids)))
    return "".join(lines)


def replicates_tab(variant_proportions_reps, cid_to_fid, bootstrap=True):
    """bootstraps.replicates.tab / jackknife.replicates.tab (traversome.py:724-736); '' without replicates."""
    if not variant_proportions_reps:
        return ""
    met = "bootstraps" if bootstrap else "jackknife"
    num_fids = len(cid_to_fid)
    lines = ["%s ID\t" % met.upper() + "\t".join("fid_%d" % (i + 1) for i in range(num_fids)) + "\n"]
    s_digit = len(str(len(variant_proportions_reps)))
    for go_r, v_prop in enumerate(variant_proportions_reps):
        line = ["%s_%0*d" % (met, s_digit, go_r)] + ["-"] * num_fids
        for cid, prop_val in v_prop.items():
            line[cid_to_fid[cid]] = "%.4f" % prop_val
        lines.append("\t".join(line) + "\n")
    return "".join(lines)


def final_result_tab(summary, variant_proportions, res_loglike, res_criterion, cid_to_fid, criterion=Criterion.BIC,
                     bootstrap=True):
    """final.result.tab (traversome.py:669-722).  `summary`: bootstrap.BootstrapSummary."""
    criterion = getattr(criterion, "value", criterion)
    met = "BOOTSTRAP" if bootstrap else "JACKKNIFE"
    num_fids = len(cid_to_fid)
    lines = ["SOLUTIONS\t%s_SUPPORT\tLOGLIKELIHOOD\t%s\t" % (met, criterion) +
             "\t".join("fid_%d" % (i + 1) for i in range(num_fids)) + "\tFREQ_STD_OF_%s\n" % met]
    unique, ordered = summary.vp_unique_results, summary.vp_unique_results_sorted
    raw_res = tuple(sorted(variant_proportions.keys(), key=lambda x: (summary.cid_sorter.get(x, 0), x)))
    num_solutions = len(unique) if raw_res in unique else len(unique) + 1
    s_digit = len(str(num_solutions))
    for go_s, chosen in enumerate(ordered):
        info = unique[chosen]
        if len(info["values"]) > 1:
            std = np.std(info["values"], axis=0, ddof=1)
            std_str = ",".join("%.4f" % s for s in std)
        else:
            std_str = ",".join("*" for _ in info["values"][0])
        line = ["%0*d\t%s\t%s\t%s" % (s_digit, go_s + 1, info["support"], info.get("raw_like", "*"),
                                      info.get("raw_criterion", "*"))] + ["-"] * num_fids + [std_str]
        raw_prop = info.get("raw_prop", "*")
        if isinstance(raw_prop, str) and raw_prop == "*":
            for cid in chosen:
                line[cid_to_fid[cid]] = "*"
        else:
            for cid, prop_val in raw_prop.items():
                line[cid_to_fid[cid]] = "%.4f" % prop_val
        lines.append("\t".join(line) + "\n")
    if not ordered or raw_res not in unique:
        head = "%0*d\t1.0\t-\t-" % (s_digit, num_solutions) if not ordered else \
            "%0*d\t0\t%s\t%s" % (s_digit, num_solutions, res_loglike, res_criterion)
        line = [head] + ["-"] * num_fids + ["-"]
        for cid, prop_val in variant_proportions.items():
            line[cid_to_fid[cid]] = "%.4f" % prop_val
        lines.append("\t".join(line) + "\n")
    return "".join(lines)


def all_tables(summary, variant_proportions, res_loglike, res_criterion, variant_paths, variant_sizes,
               be_unidentifiable_to, read_paths, variant_subpath_counters, criterion=Criterion.BIC, bootstrap=True):
    """{file name: text} of every table the ML path writes, in the reference's order of construction
    (output_variant_info first: it defines the fid numbering the other two use)."""
    cid_to_fid = cid_to_fid_table(summary.vp_unique_results_sorted, variant_proportions)
    met = "bootstraps" if bootstrap else "jackknife"
    tables = OrderedDict()
    tables["variants.info.tab"] = variants_info_tab(cid_to_fid, be_unidentifiable_to, variant_paths, variant_sizes)
    tables["readpath.information.tab"] = readpath_information_tab(cid_to_fid, read_paths, variant_paths,
                                                                  variant_subpath_counters)
    tables["readpath.record_ids.tab"] = readpath_record_ids_tab(read_paths)
    tables["final.result.tab"] = final_result_tab(summary, variant_proportions, res_loglike, res_criterion, cid_to_fid,
                                                  criterion, bootstrap)
    rep = replicates_tab(summary.variant_proportions_reps, cid_to_fid, bootstrap)
    if rep:
        tables["%s.replicates.tab" % met] = rep
    return tables


def write_tables(outdir, tables):
    for name, text in tables.items():
        with open(os.path.join(outdir, name), "w") as h:
            h.write(text)
