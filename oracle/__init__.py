"""CPU oracle for the ensemble posterior-sampling path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  The product
(``ert-conditional-diffusion-model_b200/``) never does: it fails loudly when its CUDA
library is missing instead of falling back to this code.

Contents
--------
``denoiser_oracle``  torch-CPU restatement of the reference's denoiser, schedule,
                     timestep embedding and DDPM reverse chain
                     (``ERT_Conditional_Diffusion.py:80-164``).
``stats_oracle``     numpy restatement of the ensemble statistics
                     (``ERT_Conditional_Diffusion.py:747-762, 867-872``) and of
                     the numpy / scipy arithmetic they dispatch to.
``reference_loader`` AST-extracts the reference's own functions from
                     ``/root/reference`` (exists only in the build container); used to
                     pin the restatement and to generate ``tests/golden``.
``make_golden``      the script that generated ``tests/golden/*``.

Parity status: the reference ships no tests or golden vectors, so
the restatement is pinned against outputs of the reference itself, executed
unmodified in the build container (``make_golden.py`` → ``tests/golden``).
"""
