"""Same entry point name as the reference's examples/generate_qwen2_vl.py: image + prompt -> text with the qwen2_vl application.

    python examples/generate_qwen2_vl.py --model-path /path/to/checkpoint --image cat.png [--prompt "..."] [--tp-degree N]

(thin wrapper over multimodal_demo.py, which documents the per-family processor outputs the applications take)."""
import sys

from multimodal_demo import main

if __name__ == "__main__":
    sys.argv[1:1] = ["--model-type", "qwen2_vl"]
    main()
