#! /usr/bin/env python3
"""Command-line entry point with the reference's name and flags:
    python3 processAXCTD.py -i inputaudiofile.wav -o outputASCIIfile.txt
(see axctdprocessor_b200/processAXCTD.py)."""
from axctdprocessor_b200.processAXCTD import main, parse_times, processAXCTD  # noqa: F401

if __name__ == "__main__":
    main()
