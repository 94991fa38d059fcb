"""`python -m bulletproofs_gadgets_b200.verifier <stem>` -- mirror of the reference binary src/bin/verifier.rs:46-101:
reads <stem>.gadgets/.inst/.coms/.proof, prints true / false and exits 0 / 1."""
import sys

from . import frontend


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("missing argument", file=sys.stderr)
        return 2
    ok = frontend.verifier_main(argv[0])
    print("true" if ok else "false")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
