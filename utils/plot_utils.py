"""The one helper of the reference's utils/plot_utils.py that its tests import
(perfect_repeat_finder_tests.py:5).  Plotting itself is out of scope,
so matplotlib is not imported here."""


def shift_string_by(string, shift):
    """Rotate `string` right by `shift` characters ("AGTTT" shifted by 2 -> "TTAGT");
    same contract as the reference's utils/plot_utils.py:6-9."""
    return string[-shift:] + string[:-shift]
