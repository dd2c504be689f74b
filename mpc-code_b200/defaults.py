"""Default flag values a problem file may override.

Mirror of the reference's two-stage star import (``MPC_code.py:23-28``): every name the
driver reads is pre-seeded here with the value ``Default_Values.py:16-131`` gives it,
then the user's file is executed on top.
"""

_NONE_NAMES = """umin umax xmin xmax ymin ymax umin_ss umax_ss xmin_ss xmax_ss ymin_ss ymax_ss
umin_dyn umax_dyn xmin_dyn xmax_dyn ymin_dyn ymax_dyn dmin dmax Dumin Dumax wmin wmax vmin vmax""".split()

_FALSE_NAMES = """estimating ssjacid StateFeedback Fp_nominal QForm_ss DUssForm Adaptation ContForm
TermCons QForm DUForm DUFormEcon kalss lue kal ekf mhe Collocation slacks""".split()


def default_namespace() -> dict:
    ns = {name: None for name in _NONE_NAMES}
    ns.update({name: False for name in _FALSE_NAMES})
    ns.update(
        offree="no",
        Sol_itmax=100,
        Sol_Hess_constss="no", Sol_Hess_constdyn="no", Sol_Hess_constmhe="no",
        LinPar=True, slacksG=True, slacksH=True,
    )
    return ns
