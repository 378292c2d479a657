"""Run a script of the reference checkout UNMODIFIED on the sm_100a operator surface.

    python -m imagecompression_adversarial_b200.launch --ref /path/to/ImageCompression_Adversarial \
           attack_rd.py -m hyper -metric mse -q 3 --new -steps 100 -s 'imgs/*.png'

What it does: puts the reference checkout on ``sys.path``; for every third-party module the reference imports
that is NOT installed (compressai, pytorch_msssim, lpips, thop, matplotlib) puts this package's stand-in in front
(real packages win when present -- then nothing of ours is used); runs the script as ``__main__``.
With ``--fused`` (or ICADV_FUSED=1) the reference's ``attack_rd.attack_`` is replaced by the device-resident
loop ``imagecompression_adversarial_b200.attack.attack_`` (same signature and return tuple, no per-step host sync).
"""
import importlib
import importlib.util
import os
import runpy
import sys

_SHIMS = ("compressai", "pytorch_msssim", "lpips", "thop", "matplotlib")


def install_shims():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
    missing = [m for m in _SHIMS if importlib.util.find_spec(m) is None]
    if missing:
        # one directory holds all stand-ins; shadowing an installed package is avoided by importing the installed
        # ones first so they are already in sys.modules
        for m in _SHIMS:
            if m not in missing:
                importlib.import_module(m)
        sys.path.insert(0, here)
    return missing


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    ref, fused = os.environ.get("ICADV_REFERENCE", ""), os.environ.get("ICADV_FUSED", "0") == "1"
    while argv and argv[0].startswith("--"):
        if argv[0] == "--ref":
            ref = argv[1]
            argv = argv[2:]
        elif argv[0] == "--fused":
            fused = True
            argv = argv[1:]
        else:
            break
    if not ref or not argv:
        raise SystemExit(__doc__)
    script = argv[0] if os.path.isabs(argv[0]) else os.path.join(ref, argv[0])
    missing = install_shims()
    sys.path.insert(0, ref)
    print(f"[icadv-b200] reference: {ref}; stand-ins for: {', '.join(missing) or 'none'}; fused loop: {fused}",
          file=sys.stderr)
    if fused:
        import attack_rd  # the reference module
        from imagecompression_adversarial_b200 import attack as fused_attack
        attack_rd.attack_ = fused_attack.attack_
        sys.modules["attack_rd"] = attack_rd
        sys.argv = [script] + argv[1:]
        args = attack_rd.coder.config().parse_args()
        attack_rd.args = args   # attacker.attack() reads the module-level `args` (attack_rd.py:609)
        attack_rd.main(args)
        return
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
