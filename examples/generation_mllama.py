"""Same entry point name as the reference's examples/generation_mllama.py: image + prompt -> text with the mllama application.

    python examples/generation_mllama.py --model-path /path/to/checkpoint --image cat.png [--prompt "..."] [--tp-degree N]

(thin wrapper over multimodal_demo.py, which documents the per-family processor outputs the applications take)."""
import sys

from multimodal_demo import main

if __name__ == "__main__":
    sys.argv[1:1] = ["--model-type", "mllama"]
    main()
