"""Argument validation and index selection for the host-side parameter store.

Written from the behavioural contract of the reference's setter layer (``src/_BirthDeath.pyx:1187-1702``), which its own
test-suite pins (``tests/test_interface.py``: 278 cases of stored values, exception types and exact messages).  The
messages therefore have to be the reference's, character for character; the code that produces them is organised
differently: a setter names WHAT each argument is (a `Quantity`, an `Axis` selector) and the helpers below turn that
into validated values and sorted index arrays that numpy assigns in one statement.

An axis selector is what the reference API accepts wherever it takes an index: ``None`` (everything), an int, a list of
selectors, and -- on the haplotype axis -- a string over ``ATCG`` with ``*`` wildcards (``'G*'`` = every haplotype whose
first site is G).
"""
import itertools

import numpy as np

ALPHABET = "ATCG"
_HAP_TEXT = ('Incorrect haplotype. Haplotype should contain only "A", "T", "C", "G", "*" and length of haplotype should be '
             'equal number of mutations sites.')


def _is_number(v):
    return isinstance(v, (int, float))


def _wrong_type(what, expected):
    return TypeError('Incorrect type of %s. Type should be %s.' % (what, expected))


def _wrong_value(what, rule):
    return ValueError('Incorrect value of %s. Value should be %s.' % (what, rule))


def count(value, what, positive=True):
    """A count / size / seed: an int, > 0 (or >= 0 when `positive` is False)."""
    if not isinstance(value, int):
        raise _wrong_type(what, 'int')
    if positive and value <= 0:
        raise _wrong_value(what, 'more 0')
    if not positive and value < 0:
        raise _wrong_value(what, 'more or equal 0')
    return value


def quantity(value, what, upper=None, optional=False):
    """A rate / probability / multiplier: a non-negative number, at most `upper` when given."""
    if not _is_number(value):
        if optional and value is None:
            return value
        raise _wrong_type(what, 'int or float or None' if optional else 'int or float')
    if value < 0 or (upper is not None and value > upper):
        rule = 'more or equal 0' if upper is None else 'more or equal 0 and equal or less %s' % upper
        raise _wrong_value(what, rule)
    return value


def fixed_list(data, what, length):
    if not isinstance(data, list):
        raise _wrong_type(what, 'list')
    if len(data) != length:
        raise ValueError('Incorrect length of %s. Length should be equal %d.' % (what, length))
    return data


class Axis:
    """One index dimension of the model (haplotypes, demes, susceptibility groups, mutation sites)."""

    def __init__(self, size, what, sites=None):
        self.size = size
        self.what = what
        self.sites = sites          # not None: the haplotype axis (selectors may be strings)

    # -- validation of one selector element
    def _check_one(self, sel, required):
        if sel is None:
            if required:
                raise _wrong_type(self.what, 'int')
            return
        if isinstance(sel, int):
            if not 0 <= sel < self.size:
                raise IndexError('There are no such %s!' % self.what)
            return
        if self.sites is not None:
            if isinstance(sel, str):
                if sum(sel.count(ch) for ch in ALPHABET + '*') != self.sites:
                    raise ValueError(_HAP_TEXT)
                return
            raise _wrong_type('haplotype', 'int or str or None')
        raise _wrong_type(self.what, 'int or None')

    def check(self, sel, required=False, lists=True):
        for one in (sel if (lists and isinstance(sel, list)) else [sel]):
            self._check_one(one, required)

    # -- expansion of one selector element into indices
    def _expand_one(self, sel):
        if isinstance(sel, str):
            per_site = [ALPHABET if ch == '*' else ch for ch in sel]
            return [self.encode(''.join(word)) for word in itertools.product(*per_site)]
        if isinstance(sel, int):
            return [sel]
        return range(self.size)

    def indices(self, sel, lists=True):
        """Sorted unique indices a (validated) selector stands for."""
        picked = set()
        for one in (sel if (lists and isinstance(sel, list)) else [sel]):
            picked.update(self._expand_one(one))
        return np.fromiter(sorted(picked), dtype=np.int64, count=len(picked))

    def select(self, sel, required=False, lists=True):
        self.check(sel, required=required, lists=lists)
        return self.indices(sel, lists=lists)

    # -- haplotype <-> string (site 0 is the most significant base-4 digit, A T C G = 0 1 2 3)
    def encode(self, word):
        value = 0
        for ch in word[:self.sites]:
            value = value * 4 + max(ALPHABET.find(ch), 0)
        return value

    def decode(self, value):
        word = []
        for _ in range(self.sites):
            value, digit = divmod(value, 4)
            word.append(ALPHABET[digit])
        return ''.join(reversed(word))

    def allele(self, haplotype, site):
        return (haplotype // 4 ** (self.sites - site - 1)) % 4


def close_migration_matrix(m):
    """Diagonal of the short-visit matrix = 1 - (row's off-diagonal sum); the reference's two admissibility rules.
    Rows are closed one after the other and the first inadmissible row stops the walk (rows behind it keep their old
    diagonal, as in the reference); sums run in increasing target order because the results are compared exactly."""
    K = m.shape[0]
    for p in range(K):
        leaving, stay = 0.0, 1.0
        for q in range(K):
            if q != p:
                leaving += m[p, q]
                stay -= m[p, q]
        m[p, p] = stay
        if leaving > 1:
            raise ValueError('Incorrect the sum of migration probabilities. The sum of migration probabilities from each '
                             'population should be equal or less 1.')
    if np.any(np.diagonal(m) <= 1e-15):
        raise ValueError('Incorrect value of migration probability. Value of migration probability from source population '
                         'to target population should be more 0.')
