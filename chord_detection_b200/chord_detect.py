"""`chord-detect` CLI — same flags and output lines as
/root/reference/chord_detection/chord_detect.py:11-63 (--key, --displayplots N, --method, input_path).

Batch mode: several inputs, a directory, or a manifest (*.txt / *.lst, one path
per line) run through the batched device ops (chord_detection_b200/batch.py); under torchrun the
clip list is sharded over the GPUs.  A single WAV path behaves exactly like the reference CLI."""
import argparse

from chord_detection_b200 import METHODS


def main_cli(argv=None):
    method_nums_help_string = "-1 = all, "
    for k in METHODS.keys():
        method_nums_help_string += "{0} ({1}), ".format(k, METHODS[k].display_name())
    method_nums_help_string = method_nums_help_string[:-2]

    parser = argparse.ArgumentParser(
        prog="chord-detection",
        description="Collection of chord-detection techniques",
        formatter_class=argparse.RawDescriptionHelpFormatter,
    )
    parser.add_argument(
        "--key",
        action="store_true",
        help="estimate the key using the Krumhansl-Schmuckler key-finding algorithm",
    )
    parser.add_argument(
        "--displayplots",
        type=int,
        help="frame whose intermediates are kept on the object (plots are out of scope)",
        default=-1,
    )
    parser.add_argument(
        "--method",
        type=int,
        help=method_nums_help_string,
        default=next(iter(METHODS.keys())),
    )
    parser.add_argument("--batch", action="store_true",
                        help="force batch mode for a single input (per-clip lines + corpus sums)")
    parser.add_argument("input_path", nargs="+",
                        help="Path to WAV audio clip (or several, a directory, a *.txt manifest)")
    args = parser.parse_args(argv)

    import os

    first = args.input_path[0]
    if (args.batch or len(args.input_path) > 1 or os.path.isdir(first)
            or first.lower().endswith((".txt", ".lst"))):
        from chord_detection_b200 import batch

        methods = list(METHODS.keys()) if args.method == -1 else [args.method]
        if any(m not in METHODS for m in methods):
            raise ValueError("valid methods: {0}".format(method_nums_help_string))
        batch.main_batch(args.input_path, methods, key=args.key)
        return
    args.input_path = first

    compute_objs = []
    if args.method == -1:
        for v in METHODS.values():
            compute_objs.append(v(args.input_path))
    else:
        try:
            compute_objs.append(METHODS[args.method](args.input_path))
        except KeyError:
            raise ValueError("valid methods: {0}".format(method_nums_help_string))

    for compute_obj in compute_objs:
        print("{0} - {1}".format(compute_obj.method_number(), compute_obj.display_name()))
        chromagram = compute_obj.compute_pitches(args.displayplots)
        print(chromagram)
        if args.key:
            print(chromagram.key())


if __name__ == "__main__":
    main_cli()
