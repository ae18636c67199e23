"""Host logic of the command-line mirror (axctdprocessor_b200/processAXCTD.py) that needs no device."""
import pytest

from axctdprocessor_b200 import processAXCTD as cli

# reference processAXCTD.parse_times (processAXCTD.py:106-121) evaluated on these strings in the build container
PARSE_TIMES_REFERENCE = {"0": 0, "-1": -1, "30": 30, "1:30": 90, "01:02:03": 3723, "1:2:3:4": 7384, "x:1:2:3": 3723, "a": -2,
                         "1:a": -2, "": -2, "1:": -2, "::": -2, "5:00": 300, "-5": -5, "1:-2": 58, "1.5": -2,
                         "99:99:99": 362439, "1:2:3:x": -2, "0:0:0:0:7": 7}


@pytest.mark.parametrize("text,expected", sorted(PARSE_TIMES_REFERENCE.items()))
def test_parse_times_matches_reference(text, expected):
    assert cli.parse_times(text) == expected


def test_flags_and_defaults_are_the_reference_ones():
    flags = {short: (long_name, default) for short, long_name, default, _ in cli._VALUE_FLAGS}
    assert flags == {"-i": ("--input", "ERROR_NO_FILE_SPECIFIED"), "-o": ("--output", "output.txt"), "-s": ("--starttime", "0"),
                     "-e": ("--endtime", "-1"), "-a": ("--autodetect-start", "30"), "-b": ("--autodetect-end", "-1"),
                     "-p": ("--sig-threshold-400", "2"), "-t": ("--sig-threshold-7500", "1.5"), "-d": ("--dead-freq", "3000"),
                     "-l": ("--pointsperloop", "100000"), "-m": ("--mark-freq", "400"), "-n": ("--space-freq", "800")}


def test_missing_input_prints_the_reference_message_and_exits(capsys, tmp_path):
    with pytest.raises(SystemExit):
        cli.main([])
    assert capsys.readouterr().out == "[!] Error- no input WAV file specified! Terminating\n"
    with pytest.raises(SystemExit):
        cli.main(["-i", str(tmp_path / "absent.wav")])
    assert capsys.readouterr().out == "[!] Specified input file does not exist! Terminating\n"
