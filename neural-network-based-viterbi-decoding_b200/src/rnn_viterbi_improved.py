"""Alias module: the reference's tests import `rnn_viterbi_improved` (tests/test_encoders.py:16-24,
tests/test_utilities.py:16-24), a name that does not exist in its tree; it maps to `rnn_viterbi`."""
from rnn_viterbi import *  # noqa: F401,F403
from rnn_viterbi import __doc__ as _doc  # noqa: F401
