#!/usr/bin/env python
"""Drop-in for the reference's ``code/generate_pseudo_labels.py`` (same flags; see hiast_b200/cli.py)."""
from hiast_b200.cli import main

if __name__ == '__main__':
    main()
