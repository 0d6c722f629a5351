#!/usr/bin/env python3
"""`gipuma`-compatible command line on top of libtsar_b200.so (see tsar-mvs_b200/cli.py for the flags)."""
import sys

import __graft_entry__ as g

if __name__ == "__main__":
    pkg = g.load_package()
    from tsar_mvs_b200 import cli
    sys.exit(cli.run(sys.argv[1:]))
