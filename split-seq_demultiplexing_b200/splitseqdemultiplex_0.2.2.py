#!/usr/bin/env python
"""Drop-in for the reference's python_only_tool/splitseqdemultiplex_0.2.2.py: upstream it is a stale copy of 0.2.3 (CLI3)
that crashes; here it is the same command line, served by splitseqdemultiplex_0.2.3.py next to it."""
import os
import runpy

if __name__ == '__main__':
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "splitseqdemultiplex_0.2.3.py"), run_name="__main__")
