"""The generated header - and with it the digest that names the compiled library - must not depend on what the process
traced before: the in-tree libraries are built once (`__graft_entry__.build()`), travel to the GPU box and have to be
the ones every later process asks for."""
import __graft_entry__ as entry
from mpc_code_b200.build import source_digest
from mpc_code_b200.devicegen import generate_header


def _text(name):
    prob, ss, ocp = entry._problem(name)
    return generate_header(prob, ss, ocp)["text"]


def test_header_text_is_independent_of_earlier_traces():
    first = _text("nmpc_cstr")
    _text("nmpc_cstr_gineq")            # shares the reactor's constants with nmpc_cstr
    _text("lmpc_cstr")
    again = _text("nmpc_cstr")
    assert again == first
    assert source_digest(again) == source_digest(first)
    assert _text("enmpc_reactor") == _text("enmpc_reactor")
