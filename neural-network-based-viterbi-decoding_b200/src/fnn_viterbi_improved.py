"""Alias module: the reference's tests import `fnn_viterbi_improved` (tests/test_encoders.py:16-24,
tests/test_utilities.py:16-24), a name that does not exist in its tree; it maps to `fnn_viterbi`."""
from fnn_viterbi import *  # noqa: F401,F403
from fnn_viterbi import __doc__ as _doc  # noqa: F401
