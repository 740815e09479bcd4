"""Same entry point name as the reference's examples/generation_llama4.py: image + prompt -> text with the llama4 application.

    python examples/generation_llama4.py --model-path /path/to/checkpoint --image cat.png [--prompt "..."] [--tp-degree N]

(thin wrapper over multimodal_demo.py, which documents the per-family processor outputs the applications take)."""
import sys

from multimodal_demo import main

if __name__ == "__main__":
    sys.argv[1:1] = ["--model-type", "llama4"]
    main()
