"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of Traversome's ML frequency-fit path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker or the timed CPU baseline -- never as the
product path.  ``traversome_b200`` must not import it (tests/test_no_oracle_in_product.py enforces it).

Pinning status: the reference ships no tests, fixtures or golden vectors
for this path.  The oracle is therefore pinned against outputs of the reference ITSELF, generated
in the build container by ``oracle/make_golden.py`` and committed under ``tests/golden/``:
  * log-likelihood known answers: the reference's own ``PathMultinomialModel.get_like_formula``
    evaluated with floats and ``math.log`` (no substitution involved) -- matched bit for bit;
  * integer tables (from_variants, num_matched, variant_sizes): the reference's own -- matched exactly;
  * fitted proportions / model-selection results: "reference + sympy stand-in for symengine"
    (symengine is not installable here), matched within the reference's own SLSQP tolerance.
"""
