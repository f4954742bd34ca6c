"""CPU oracle for the MagmaClustR hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/LAPACK restatement of the reference's algorithm
(ArthurLeroy/MagmaClustR, R sources cited file:line in every function).  It is
the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``magmaclustr_b200``) never imports, links or executes anything here.

Parity status (see DESIGN.md "Oracle"):
  * kernels / kern_to_cov / kern_to_inv / gradients: pinned by the reference's
    own known-answer and finite-difference tests (tests/test_oracle_*.py).
  * e_step, m_step, logL_monitoring, train_magma, ve_step, vm_step,
    update_mixture, ELBOs, hyperposterior, pred_magma*: **parity unpinned** --
    the reference holds no numeric fixture for them and R cannot run in this
    image.  The L-BFGS-B optimiser is a restatement of the Netlib L-BFGS-B 2.x
    code R ships (src/appl/lbfgsb.c), also unpinned.
"""
