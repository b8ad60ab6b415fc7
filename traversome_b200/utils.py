"""Record types and criteria of the ML-fit boundary.

Host-side mirror of the few names of ``traversome/utils.py`` that cross the boundary of the hot
path: ``Bins`` / ``BinInfo`` / ``SubPathInfo`` (utils.py:193-226),
``LogLikeFormulaInfo`` / ``LogLikeFuncInfo`` (utils.py:229-240), ``Criterion`` (utils.py:243-245)
and ``aic`` / ``bic`` (utils.py:248-253).  Same attribute names and meaning, so reference objects
and these objects are interchangeable ("duck-typed") everywhere in this package.
"""
from enum import Enum
from math import log


class SubPathInfo(object):
    """One read path (= a CSR row): where it occurs and which records support it (utils.py:193-203)."""
    __slots__ = ("full_len", "inner_len", "from_variants", "mapped_records", "num_possible_X", "num_in_range")

    def __init__(self):
        self.full_len = None
        self.inner_len = None
        self.from_variants = {}      # {variant id: occurrences of this read path in the variant}
        self.mapped_records = []     # alignment record ids
        self.num_possible_X = None
        self.num_in_range = None


class BinInfo(object):
    """One multinomial bin = (alignment-length range, read path) (utils.py:220-224)."""
    __slots__ = ("from_variants", "num_possible_X", "num_matched")

    def __init__(self, from_variants=None, num_possible_X=None, num_matched=0):
        self.from_variants = {} if from_variants is None else from_variants
        self.num_possible_X = num_possible_X   # mean number of possible alignment start points
        self.num_matched = num_matched         # observed records


class Bins(object):
    """All bins of one alignment-length range (utils.py:206-217)."""
    __slots__ = ("rp_bins", "max_len", "min_len", "min_id", "max_id")

    def __init__(self, min_len=None, max_len=None, min_id=None, max_id=None, rp_bins=None):
        self.rp_bins = [] if rp_bins is None else rp_bins
        self.max_len = max_len
        self.min_len = min_len
        self.min_id = min_id
        self.max_id = max_id


class LogLikeFormulaInfo(object):
    __slots__ = ("loglike_expression", "variable_size", "sample_size")

    def __init__(self, loglike_expression=0, variable_size=0, sample_size=0):
        self.loglike_expression = loglike_expression
        self.variable_size = variable_size
        self.sample_size = sample_size


class LogLikeFuncInfo(object):
    __slots__ = ("loglike_func", "variable_size", "sample_size")

    def __init__(self, loglike_func=0, variable_size=0, sample_size=0):
        self.loglike_func = loglike_func
        self.variable_size = variable_size
        self.sample_size = sample_size


class Criterion(str, Enum):
    AIC = "AIC"
    BIC = "BIC"


def aic(loglike, len_param):
    """Akaike information criterion, 2k - 2LL (utils.py:252-253)."""
    return 2 * len_param - 2 * loglike


def bic(loglike, len_param, len_data):
    """Bayesian information criterion, ln(N) k - 2LL (utils.py:248-249)."""
    return log(len_data) * len_param - 2 * loglike
