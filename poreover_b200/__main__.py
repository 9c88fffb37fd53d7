"""`python -m poreover_b200 {decode,pair-decode}`: the reference's command line (poreover/__main__.py:53-91),
same flags and defaults, backed by the B200 kernels.  train / call / benchmark are not part of the decoding
hot path and are not provided."""
import argparse
import sys

from . import __version__


def build_parser():
    parser = argparse.ArgumentParser(prog="poreover_b200", description='PoreOver decoding on B200')
    sub = parser.add_subparsers(dest="command")
    sub.required = True
    d = sub.add_parser('decode', help='Decode basecaller probabilities to a FASTA file')
    d.add_argument('in', nargs='+')
    d.add_argument('--out', default='out')
    d.add_argument('--basecaller', choices=['poreover', 'flappie', 'guppy', 'bonito'])
    d.add_argument('--algorithm', default='viterbi', choices=['viterbi', 'beam', 'prefix'])
    d.add_argument('--window', type=int, default=400)
    d.add_argument('--beam_width', type=int, default=25)
    d.add_argument('--threads', type=int, default=1, help='accepted for compatibility; batching replaces the pool')
    d.add_argument('-v', '--version', action='version', version=__version__)
    p = sub.add_parser('pair-decode', help='1D2 consensus decoding of two output probabilities',
                       formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('-v', '--version', action='version', version=__version__)
    p.add_argument('in', nargs='+')
    p.add_argument('--dir', default='.')
    p.add_argument('--basecaller', choices=['poreover', 'flappie', 'guppy', 'bonito'])
    p.add_argument('--reverse_complement', default=False, action='store_true')
    p.add_argument('--out', default='out')
    p.add_argument('--threads', type=int, default=1, help='accepted for compatibility; batching replaces the pool')
    p.add_argument('--method', choices=['align', 'split', 'envelope'], default='envelope', help=argparse.SUPPRESS)
    p.add_argument('--single', choices=['beam', 'viterbi'], default='viterbi')
    p.add_argument('--logging', default="info", choices=['info', 'debug'])
    p.add_argument('--debug', default=False, action='store_true')
    p.add_argument('--algorithm', default='beam', choices=['prefix', 'beam'], help=argparse.SUPPRESS)
    p.add_argument('--alignment', default='banded', choices=['banded', 'full'])
    p.add_argument('--beam_width', type=int, default=5)
    p.add_argument('--debug_envelope', action='store_true', help=argparse.SUPPRESS)
    p.add_argument('--diagonal_envelope', action='store_true')
    p.add_argument('--diagonal_width', type=int, default=50)
    p.add_argument('--padding', type=int, default=5)
    p.add_argument('--skip_matches', action='store_true')
    p.add_argument('--skip_threshold', type=int, default=10)
    p.add_argument('--beam_search_method', choices=['row', 'row_col', 'grid'], default="row_col", help=argparse.SUPPRESS)
    p.add_argument('--window', type=int, default=200, help=argparse.SUPPRESS)
    return parser


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    # the reference takes any width; the GPU searches run one thread per (node, read) of the expanded beam
    uses_beam = args.command == 'pair-decode' or getattr(args, 'algorithm', None) == 'beam'
    if uses_beam and not 1 <= args.beam_width <= 100:
        parser.error("--beam_width must be between 1 and 100 on the B200 backend (got %d)" % args.beam_width)
    if args.command == 'decode':
        from .decoding.decode import decode
        decode(args)
    else:
        from .decoding.pair_decode import pair_decode
        pair_decode(args)
    print(args, file=sys.stderr)


if __name__ == "__main__":
    main()
