"""`python -m bulletproofs_gadgets_b200.prover <stem>` -- mirror of the reference binary src/bin/prover.rs:47-100:
reads <stem>.gadgets/.inst/.wtns, writes <stem>.coms and <stem>.proof, prints the number of constraints."""
import sys

from . import frontend


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("missing argument", file=sys.stderr)
        return 2
    print(frontend.prover_main(argv[0]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
