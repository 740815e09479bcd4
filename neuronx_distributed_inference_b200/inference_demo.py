"""``inference_demo`` CLI (reference inference_demo.py:73-799):

    inference_demo --model-type llama --task-type causal-lm run --model-path ... --compiled-model-path ... \\
        --torch-dtype bfloat16 --tp-degree 8 --batch-size 2 --max-context-length 32 --seq-len 64 \\
        --on-device-sampling --enable-bucketing --top-k 1 --pad-token-id 2 --prompt "..." \\
        --check-accuracy-mode token-matching --benchmark

``--kebab-flags`` map 1:1 to NeuronConfig kwargs (``None`` dropped, reference :441-442).  With ``--tp-degree N > 1``
the CLI re-launches itself under ``torch.distributed.run`` (one process per GPU) unless already inside a job.
Neuron-only flags (LNC, compiler, DGE, cc tiling, NKI kernel toggles ...) are accepted and ignored with a warning.
"""
from __future__ import annotations

import argparse
import ast
import copy
import json
import logging
import os
import subprocess
import sys
from enum import Enum

import torch

from .config import FusedSpecNeuronConfig, OnDeviceSamplingConfig, to_torch_dtype, load_pretrained_config
from .utils.constants import MODEL_TYPES, TASK_TYPES, get_model_cls

logger = logging.getLogger("b200infer")


class CheckAccuracyMode(Enum):
    SKIP_ACCURACY_CHECK = "skip-accuracy-check"
    TOKEN_MATCHING = "token-matching"
    LOGIT_MATCHING = "logit-matching"
    DRAFT_LOGIT_MATCHING = "draft-logit-matching"

    def __str__(self):
        return self.value


# (flag, kwargs)  — every entry becomes a NeuronConfig kwarg named after the flag
_INT, _STR, _FLOAT = dict(type=int), dict(type=str), dict(type=float)
_FLAG = dict(action="store_true", default=None)
_DTYPE = dict(type=to_torch_dtype)
CONFIG_FLAGS = [
    # basic (reference :126-141)
    ("torch-dtype", _DTYPE), ("batch-size", _INT), ("padding-side", _STR), ("allow-input-truncation", _FLAG),
    ("seq-len", _INT), ("n-active-tokens", _INT), ("n-positions", _INT), ("max-context-length", _INT),
    ("max-new-tokens", _INT), ("max-length", _INT), ("rpl-reduce-dtype", _DTYPE), ("attention-dtype", _DTYPE),
    ("output-logits", _FLAG), ("vocab-parallel", _FLAG), ("layer-boundary-markers", _FLAG),
    # attention (:143-147)
    ("fused-qkv", _FLAG), ("sequence-parallel-enabled", _FLAG), ("weight-gather-seq-len-threshold", _INT),
    ("flash-decoding-enabled", _FLAG), ("rolling-sliding-window-cache", _FLAG),
    # continuous batching (:149-153)
    ("ctx-batch-size", _INT), ("tkg-batch-size", _INT), ("max-batch-size", _INT), ("is-continuous-batching", _FLAG),
    # KV (:155-158)
    ("kv-cache-batch-size", _INT), ("kv-cache-padding-size", _INT), ("disable-kv-cache-tiling", _FLAG),
    # bucketing (:164-169)
    ("enable-bucketing", _FLAG), ("bucket-n-active-tokens", _FLAG),
    ("context-encoding-buckets", dict(nargs="+", type=int)), ("prefix-buckets", dict(nargs="+", type=int)),
    ("token-generation-buckets", dict(nargs="+", type=int)), ("token-generation-batches", dict(nargs="+", type=int)),
    # quantization (:171-201)
    ("quantized", _FLAG), ("quantized-checkpoints-path", _STR), ("quantization-type", _STR), ("quantization-dtype", _STR),
    ("kv-cache-quant", _FLAG), ("quantized-mlp-kernel-enabled", _FLAG), ("activation-quantization-type", _STR),
    ("rmsnorm-quantize-kernel-enabled", _FLAG), ("quantize-clamp-bound", _FLOAT), ("quantization-block-size", dict(nargs="+", type=int)),
    # MoE (:203-211)
    ("capacity-factor", _FLOAT), ("moe-tp-degree", _INT), ("moe-ep-degree", _INT), ("early-expert-affinity-modulation", _FLAG),
    ("disable-normalize-top-k-affinities", _FLAG), ("fused-shared-experts", _FLAG), ("return-expert-index", _FLAG),
    # speculation (:213-231)
    ("speculation-length", _INT), ("spec-batch-size", _INT), ("enable-fused-speculation", _FLAG),
    ("enable-eagle-speculation", _FLAG), ("enable-eagle-draft-input-norm", _FLAG), ("is-eagle3", _FLAG),
    ("is-medusa", _FLAG), ("medusa-speculation-length", _INT), ("num-medusa-heads", _INT),
    # parallelism (:233-261)
    ("tp-degree", _INT), ("cp-degree", _INT), ("attention-dp-degree", _INT), ("pp-degree", _INT), ("ep-degree", _INT),
    ("world-size", _INT), ("start-rank-id", _INT), ("local-ranks-size", _INT), ("save-sharded-checkpoint", _FLAG),
    ("skip-sharding", _FLAG), ("mlp-cp-degree", _INT),
    # paged attention / prefix caching (:263-275)
    ("is-block-kv-layout", _FLAG), ("pa-num-blocks", _INT), ("pa-block-size", _INT), ("is-prefix-caching", _FLAG),
    # async / windowed CTE (:277-281)
    ("async-mode", _FLAG), ("windowed-context-encoding-size", _INT),
    # kernels / compiler: accepted, no effect on B200 (:299-334)
    ("attn-kernel-enabled", _FLAG), ("qkv-kernel-enabled", _FLAG), ("mlp-kernel-enabled", _FLAG),
    ("attn-block-tkg-nki-kernel-enabled", _FLAG), ("attn-tkg-nki-kernel-enabled", _FLAG),
    ("attn-tkg-builtin-kernel-enabled", _FLAG), ("qkv-kernel-nbsd-layout", _FLAG), ("mlp-kernel-fuse-residual-add", _FLAG),
    ("qkv-kernel-fuse-residual-add", _FLAG), ("fused-rmsnorm-skip-gamma", _FLAG), ("out-proj-kernel-enabled", _FLAG),
    ("k-cache-transposed", _FLAG), ("logical-nc-config", _INT), ("cc-pipeline-tiling-factor", _INT),
    ("enable-spill-reload-dge", _FLAG), ("scratchpad-page-size", _INT), ("target", _STR),
    ("attn-block-tkg-nki-kernel-cache-update", _FLAG), ("attn-block-tkg-nki-kernel-cascaded-attention", _FLAG),
    ("qkv-nki-kernel-enabled", _FLAG), ("qkv-cte-nki-kernel-fuse-rope", _FLAG), ("strided-context-parallel-kernel-enabled", _FLAG),
    ("enable-cte-modular-flow", _FLAG), ("enable-output-completion-notifications", _FLAG), ("logical-neuron-cores", _INT),
    # run control (:360-408)
    ("skip-warmup", _FLAG), ("apply-seq-ids-mask", _FLAG), ("on-cpu", _FLAG),
    ("cast-type", dict(choices=["config", "as-declared"], default=None)),
    # B200-native
    ("cuda-graphs", dict(action=argparse.BooleanOptionalAction, default=None)),
    ("fused-collectives", dict(action=argparse.BooleanOptionalAction, default=None)),
]


def setup_run_parser(p: argparse.ArgumentParser):
    p.add_argument("--model-path", type=str, required=True)
    p.add_argument("--compiled-model-path", type=str, required=True)
    # evaluation (:103-124)
    p.add_argument("--benchmark", action="store_true")
    p.add_argument("--check-accuracy-mode", type=CheckAccuracyMode, choices=list(CheckAccuracyMode),
                   default=CheckAccuracyMode.SKIP_ACCURACY_CHECK)
    p.add_argument("--expected-outputs-path", type=str)
    p.add_argument("--divergence-difference-tol", type=float, default=0.001)
    p.add_argument("--tol-map", type=str)
    p.add_argument("--num-tokens-to-check", type=int)
    p.add_argument("--prompt", dest="prompts", type=str, action="append", required=True)
    p.add_argument("--top-k", type=int, default=1)
    p.add_argument("--top-p", type=float, default=1.0)
    p.add_argument("--temperature", type=float, default=1.0)
    p.add_argument("--global-topk", type=int)
    p.add_argument("--do-sample", action="store_true", default=False)
    p.add_argument("--dynamic", action="store_true", default=False)
    p.add_argument("--pad-token-id", type=int, default=0)
    p.add_argument("--top-k-kernel-enabled", action="store_true", default=False)
    p.add_argument("--on-device-sampling", action="store_true")
    p.add_argument("--sampling-dp-degree", type=int)
    for flag, kw in CONFIG_FLAGS:
        p.add_argument("--" + flag, **kw)
    # spellings of the reference that map onto the same config fields (reference :243-275)
    p.add_argument("--start_rank_id", dest="start_rank_id", type=int)
    p.add_argument("--local_ranks_size", dest="local_ranks_size", type=int)
    p.add_argument("--enable-block-kv-layout", dest="is_block_kv_layout", action="store_true", default=None)
    p.add_argument("--enable-prefix-caching", dest="is_prefix_caching", action="store_true", default=None)
    p.add_argument("--enable-chunked-prefill", dest="is_chunked_prefill", action="store_true", default=None)
    p.add_argument("--enable-torch-dist", action="store_true",
                   help="initialise torch.distributed (gloo) even for a single-rank CPU run (multi-node examples of the reference)")
    # KV-cache quantisation (reference :186-195)
    p.add_argument("--k-quant-method", type=str, default="per_tensor_symmetric",
                   help="per_tensor_symmetric | per_channel_symmetric | per_key_symmetric")
    p.add_argument("--v-quant-method", type=str, default="per_tensor_symmetric")
    p.add_argument("--kv-quant-dtype", type=str, default="float8_e4m3fn")
    p.add_argument("--kv-direct-cast", action=argparse.BooleanOptionalAction, default=True,
                   help="direct cast to the KV dtype (no scales); --no-kv-direct-cast uses the k/v quantisation methods")
    # MoE router (reference :210-211)
    p.add_argument("--router-act-fn", type=str)
    p.add_argument("--router-dtype", type=str)
    # speculation / draft
    p.add_argument("--draft-model-path", type=str)
    p.add_argument("--compiled-draft-model-path", type=str)
    p.add_argument("--draft-model-tp-degree", type=int)
    p.add_argument("--no-trim-draft-model", action="store_true")
    p.add_argument("--medusa-tree-json", type=str)
    p.add_argument("--token-tree-json", type=str)
    # LoRA (:287-297)
    p.add_argument("--enable-lora", action="store_true")
    p.add_argument("--max-loras", type=int, default=1)
    p.add_argument("--max-lora-rank", type=int, default=16)
    p.add_argument("--max-cpu-loras", type=int, default=0)
    p.add_argument("--target-modules", nargs="+")
    p.add_argument("--lora-ckpt-path", dest="lora_ckpt_paths", type=str, action="append")
    p.add_argument("--lora-ckpt-path-cpu", dest="lora_ckpt_paths_cpu", type=str, action="append")
    p.add_argument("--enable-dynamic-multi-lora", action="store_true")
    p.add_argument("--lora-ckpt-json", dest="lora_ckpt_json", type=str, default=None,
                   help='JSON file {"lora-ckpt-dir": ..., "lora-ckpt-paths": {id: path}, "lora-ckpt-paths-cpu": {id: path}}')
    p.add_argument("--adapter-id", dest="adapter_ids", type=str, action="append")
    p.add_argument("--modules-to-not-convert-file", type=str)
    # report / debug (:339-358)
    p.add_argument("--benchmark-report-path", type=str, default="./benchmark_report.json")
    p.add_argument("--num-runs", type=int, default=20)
    p.add_argument("--capture-indices", nargs="+", type=int)
    p.add_argument("--input-capture-save-dir", type=str)
    p.add_argument("--input-start-offsets", nargs="+", type=int)
    p.add_argument("--skip-compile", action="store_true")
    p.add_argument("--compile-only", action="store_true")
    p.add_argument("--compile-dry-run", action="store_true")
    p.add_argument("--hlo-debug", action="store_true")
    p.add_argument("--max-num-seqs", type=int)
    p.add_argument("--no-launch", action="store_true", help="do not re-launch under torchrun for tp-degree > 1")


def parse_args(argv=None):
    parser = argparse.ArgumentParser(prog="inference_demo")
    parser.add_argument("--model-type", type=str, choices=sorted(MODEL_TYPES), required=True)
    parser.add_argument("--task-type", type=str, choices=TASK_TYPES, required=True)
    sub = parser.add_subparsers(dest="command", required=True)
    setup_run_parser(sub.add_parser("run"))
    return parser.parse_args(argv)


def create_neuron_config(model_cls, args):
    """args -> NeuronConfig kwargs (reference :436-490)."""
    kw = {}
    for flag, _ in CONFIG_FLAGS:
        v = getattr(args, flag.replace("-", "_"), None)
        if v is not None:
            kw[flag.replace("-", "_")] = v
    kw["pad_token_id"] = args.pad_token_id
    if args.on_device_sampling:
        ods = dict(do_sample=args.do_sample, top_k=args.top_k, top_p=args.top_p, temperature=args.temperature,
                   dynamic=args.dynamic)
        if args.global_topk is not None:
            ods["global_topk"] = args.global_topk
        if args.sampling_dp_degree is not None:
            ods["sampling_dp_degree"] = args.sampling_dp_degree
        kw["on_device_sampling_config"] = OnDeviceSamplingConfig(**ods)
    if args.enable_lora:
        from .config import LoraServingConfig
        def kv(items):
            return dict(i.split(":", 1) if ":" in i else (os.path.basename(i), i) for i in (items or []))
        paths, paths_cpu = kv(args.lora_ckpt_paths), kv(args.lora_ckpt_paths_cpu)
        if args.lora_ckpt_json:      # reference lora_serving/config.py: one JSON file naming every adapter (HBM-resident and host-resident)
            with open(args.lora_ckpt_json) as f:
                j = json.load(f)
            base = j.get("lora-ckpt-dir", "")
            paths.update({k: os.path.join(base, v) for k, v in (j.get("lora-ckpt-paths") or {}).items()})
            paths_cpu.update({k: os.path.join(base, v) for k, v in (j.get("lora-ckpt-paths-cpu") or {}).items()})
        max_cpu = args.max_cpu_loras
        if args.enable_dynamic_multi_lora and max_cpu <= 0:
            max_cpu = max(len(paths_cpu), args.max_loras)      # dynamic multi-LoRA = adapters swapped in from host memory
        kw["lora_config"] = LoraServingConfig(max_loras=args.max_loras, max_lora_rank=args.max_lora_rank,
                                              max_cpu_loras=max_cpu, target_modules=args.target_modules,
                                              lora_ckpt_paths=paths, lora_ckpt_paths_cpu=paths_cpu)
    for f_ in ("is_block_kv_layout", "is_prefix_caching"):
        if getattr(args, f_, None):
            kw[f_] = True
    if kw.get("kv_cache_quant"):
        from .config import KVQuantizationConfig
        direct = args.kv_direct_cast
        # per_key_symmetric == one scale per KV head ("per_head"); per_tensor / per_channel as named
        mode = {"per_tensor_symmetric": "per_tensor", "per_channel_symmetric": "per_channel", "per_key_symmetric": "per_head"}
        if not direct and args.k_quant_method != args.v_quant_method:
            raise ValueError("--k-quant-method and --v-quant-method must agree (one scale mode per cache)")
        if not direct and args.k_quant_method not in mode:
            raise ValueError(f"unknown KV quantisation method {args.k_quant_method!r}")
        kw["kv_quant_config"] = KVQuantizationConfig(dtype=str(args.kv_quant_dtype).replace("torch.", ""),
                                                     scale_mode="direct_cast" if direct else mode[args.k_quant_method])
    if args.router_act_fn or args.router_dtype:
        kw["router_config"] = {"act_fn": args.router_act_fn or "softmax", "dtype": args.router_dtype or "float32"}
    if args.modules_to_not_convert_file:
        with open(args.modules_to_not_convert_file) as f:
            d = json.load(f)
        kw["modules_to_not_convert"] = d.get("model", d) if isinstance(d, dict) else d
        if isinstance(d, dict) and "draft_model" in d:
            kw["draft_model_modules_to_not_convert"] = d["draft_model"]
    if args.token_tree_json:
        with open(args.token_tree_json) as f:
            kw["token_tree_config"] = json.load(f)
    if args.medusa_tree_json:
        with open(args.medusa_tree_json) as f:
            kw["medusa_tree"] = json.load(f)
    if args.max_num_seqs or getattr(args, "is_chunked_prefill", None):
        from .config import ChunkedPrefillConfig
        kw["chunked_prefill_config"] = ChunkedPrefillConfig(**({"max_num_seqs": args.max_num_seqs} if args.max_num_seqs else {}))
        kw.setdefault("is_block_kv_layout", True)      # chunked prefill runs on the paged cache (config.py validation)
    return model_cls.get_neuron_config_cls()(**kw)


def _maybe_relaunch(args, argv):
    tp = args.tp_degree or 1
    if tp <= 1 or args.on_cpu and tp <= 1 or "RANK" in os.environ or args.no_launch:
        return False
    n = tp          # experts are sharded INSIDE the tensor-parallel world (ep x moe_tp == tp): no extra processes
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", os.environ.get("MASTER_PORT", "29533"), "-m", "neuronx_distributed_inference_b200.inference_demo"] + list(argv)
    logger.info("launching %d ranks: %s", n, " ".join(cmd))
    rc = subprocess.call(cmd)
    sys.exit(rc)


def run_inference(model_cls, args):
    """compile -> load -> accuracy -> generate -> benchmark (reference :493-668)."""
    from transformers import AutoTokenizer, GenerationConfig
    if getattr(args, "enable_torch_dist", False):
        # reference :245-250 — keep the processes of a multi-node example in step through torch.distributed even when the model
        # itself is single-rank; rendezvous from the usual MASTER_ADDR / RANK / WORLD_SIZE environment (defaults: one local rank)
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29534")
            dist.init_process_group("gloo" if args.on_cpu or not torch.cuda.is_available() else "nccl",
                                    rank=int(os.environ.get("RANK", "0")), world_size=int(os.environ.get("WORLD_SIZE", "1")))
    nc = create_neuron_config(model_cls, args)
    cfg_cls = model_cls.get_config_cls()
    config = cfg_cls(nc, load_config=load_pretrained_config(args.model_path))
    draft_model = None
    if args.draft_model_path is not None and not nc.enable_fused_speculation:
        dnc = copy.deepcopy(nc)
        dnc.speculation_length = 0
        dnc.is_draft_model = True
        if args.draft_model_tp_degree:
            dnc.tp_degree = args.draft_model_tp_degree
        dcfg = cfg_cls(dnc, load_config=load_pretrained_config(args.draft_model_path))
        draft_model = model_cls(args.draft_model_path, dcfg)
    elif args.draft_model_path is not None:
        dnc = copy.deepcopy(nc)
        dnc.enable_fused_speculation = False
        dnc.speculation_length = 0
        dnc.is_draft_model = True
        dnc.is_eagle_draft = bool(nc.enable_eagle_speculation)   # EAGLE: the draft consumes target features
        dnc.enable_eagle_speculation = False
        dnc.token_tree_config, dnc.enable_token_tree, dnc.is_medusa = None, False, False
        dcfg = cfg_cls(dnc, load_config=load_pretrained_config(args.draft_model_path))
        config.fused_spec_config = FusedSpecNeuronConfig(model_cls._model_cls, draft_config=dcfg,
                                                         draft_model_path=args.draft_model_path)
    model = model_cls(args.model_path, config)
    rank = int(os.environ.get("RANK", "0"))
    if nc.quantized and nc.quantized_checkpoints_path and not os.path.exists(nc.quantized_checkpoints_path):
        if rank == 0:
            model_cls.save_quantized_state_dict(args.model_path, config)   # "# Quantize model." (reference :541-543)
        if torch.distributed.is_initialized():
            torch.distributed.barrier()
    if not args.skip_compile:
        model.compile(args.compiled_model_path, dry_run=args.compile_dry_run)
        if draft_model is not None and args.compiled_draft_model_path:
            draft_model.compile(args.compiled_draft_model_path)
    if args.compile_only or args.compile_dry_run:
        return model
    model.load(args.compiled_model_path)
    if draft_model is not None:
        draft_model.load(args.compiled_draft_model_path)
    # debug hooks (reference :616-649 and application_base.py:423-554)
    if args.capture_indices:
        from .utils.debug_utils import capture_model_inputs
        capture_model_inputs(model, args.capture_indices, args.input_capture_save_dir or "saved_inputs")
    from .utils.snapshot import maybe_register_from_env
    maybe_register_from_env(model)          # NXD_INFERENCE_CAPTURE_SNAPSHOT=1 ...
    tokenizer = AutoTokenizer.from_pretrained(args.model_path, padding_side=nc.padding_side)
    if tokenizer.pad_token_id is None:
        tokenizer.pad_token_id = args.pad_token_id
    if rank == 0 and args.compiled_model_path:
        tokenizer.save_pretrained(args.compiled_model_path)
    gen_kw = dict(do_sample=args.do_sample, top_k=args.top_k, top_p=args.top_p, temperature=args.temperature,
                  pad_token_id=args.pad_token_id)
    try:
        gc = GenerationConfig.from_pretrained(args.model_path)
        gc.update(**gen_kw)
    except Exception:
        gc = GenerationConfig(**gen_kw)
    run_accuracy_check(model, tokenizer, gc, args, draft_model)
    run_generation(model, tokenizer, args.prompts, gc, draft_model, rank)
    if args.benchmark:
        from .utils.benchmark import benchmark_sampling
        rep = benchmark_sampling(model, draft_model, gc, num_runs=args.num_runs,
                                 benchmark_report_path=args.benchmark_report_path if rank == 0 else None)
        if rank == 0:
            print("Benchmark completed and its result is as following")
            print(json.dumps(rep, indent=4))
    return model


def run_accuracy_check(model, tokenizer, gc, args, draft_model=None):
    from .utils import accuracy
    mode = args.check_accuracy_mode
    if mode == CheckAccuracyMode.SKIP_ACCURACY_CHECK:
        return
    expected = torch.load(args.expected_outputs_path) if args.expected_outputs_path else None
    if mode == CheckAccuracyMode.TOKEN_MATCHING:
        accuracy.check_accuracy(model, tokenizer, gc, expected_token_ids=expected, prompts=args.prompts,
                                num_tokens_to_check=args.num_tokens_to_check, assistant_model=draft_model)
    elif mode == CheckAccuracyMode.LOGIT_MATCHING:
        tol = ast.literal_eval(args.tol_map) if args.tol_map else None
        accuracy.check_accuracy_logits(model, tokenizer, gc, expected_logits=expected, prompts=args.prompts,
                                       divergence_difference_tol=args.divergence_difference_tol, tol_map=tol,
                                       num_tokens_to_check=args.num_tokens_to_check)
    elif mode == CheckAccuracyMode.DRAFT_LOGIT_MATCHING:
        # fused speculation: the draft's logits against golden draft logits (reference utils/accuracy.py:1222-1277)
        assert expected is not None, "--expected-outputs-path with golden draft logits is required"
        ids = tokenizer(args.prompts, padding=True, return_tensors="pt")
        draft = getattr(model, "draft_model", None)
        assert draft is not None, "draft-logit-matching needs a fused-speculation model"
        pos = (ids.attention_mask.long().cumsum(-1) - 1).clamp_min(0).to(torch.int32)
        out = draft(ids.input_ids.to(model.device), ids.attention_mask.to(model.device), pos.to(model.device), None, None,
                    is_prefill=True, all_positions=True, output_logits=True)
        tol = ast.literal_eval(args.tol_map) if args.tol_map else None
        accuracy.check_draft_logits(out.logits.float().cpu().transpose(0, 1), expected, tol_map=tol)      # [steps, B, V]
    else:
        raise NotImplementedError(f"accuracy mode {mode}")
    print("Accuracy check passed")


def run_generation(model, tokenizer, prompts, gc, draft_model=None, rank=0):
    from .utils.accuracy import get_generate_outputs
    _, texts = get_generate_outputs(model, prompts, tokenizer, generation_config=gc,
                                    max_length=model.neuron_config.max_length, assistant_model=draft_model)
    if rank == 0:
        print("Generated outputs:")
        for i, t in enumerate(texts):
            print(f"Output {i}: {t}")
    return texts


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    logging.basicConfig(level=logging.INFO)
    args = parse_args(argv)
    _maybe_relaunch(args, argv)
    model_cls = get_model_cls(args.model_type, args.task_type)
    run_inference(model_cls, args)
    if torch.distributed.is_initialized():
        torch.cuda.synchronize() if torch.cuda.is_available() else None
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
