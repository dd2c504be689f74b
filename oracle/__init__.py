"""CPU oracle for the batched MPC step - TEST INFRASTRUCTURE ONLY.

Nothing in the product package may import, call or link this directory; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

It restates, on the CPU in float64 (NumPy/SciPy + generated C for the model maps), what the
reference computes per closed-loop step: the estimator update (``Estimator.py:231-386``), the
target problem (``Target_Calc.py:20-161``), the dynamic OCP (``Control_Calc.py:20-260``) and
the loop glue (``MPC_code.py:485-827``), with a dense primal-dual interior-point method that
follows IPOPT's published algorithm (Waechter & Biegler 2006) in place of the un-vendored
CasADi/IPOPT/MUMPS stack.

PARITY UNPINNED: the reference ships no tests or golden vectors and CasADi/IPOPT cannot be
installed here, so the oracle is pinned only by the known-answer facts KAT1-KAT3 derivable from
the reference's example constants and by independent SciPy cross-checks.
"""
