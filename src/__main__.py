"""`python3 src` (reference README.md:7-16): runs cellcomm_b200's entry point."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from cellcomm_b200.__main__ import main  # noqa: E402

if __name__ == '__main__':
    main(sys.argv)
