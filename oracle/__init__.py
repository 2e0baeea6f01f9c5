"""CPU oracle for the sparse-NNLS coreset hot path.  TEST INFRASTRUCTURE ONLY.

This package is a float64 NumPy restatement of the reference algorithms
(trevorcampbell/bayesian-coresets) that sit on the accelerated path:

  oracle.greedy    -- SparseNNLS greedy loop + GIGA / Frank-Wolfe / OrthoPursuit
                      (bayesiancoresets/snnls/{snnls,giga,frankwolfe,orthopursuit}.py)
  oracle.models    -- log-likelihoods of the LR / Gaussian / Poisson example models and the
                      row-centring black-box projection
                      (bayesiancoresets/projector.py, examples/common/model_{lr,gaussian,poiss}.py)
  oracle.coresets  -- Hilbert / SparseVI / BatchPSVI wiring + nn_opt
                      (bayesiancoresets/coreset/{hilbert,sparsevi,bpsvi}.py, util/opt.py)

Parity status: PINNED.  The reference publishes no tests or golden vectors, so the oracle is
pinned against outputs of the unmodified reference itself: `oracle/make_golden.py` imports the
reference from /root/reference, runs it on seeded inputs and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this restatement bit-for-bit (float64) against those
fixtures and against the known answers recorded in SURVEY.md section 8c.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package, and only as the checker / CPU baseline.  The product
package (`bayesian-coresets_b200/`) never imports it and has no CPU compute path.
"""
