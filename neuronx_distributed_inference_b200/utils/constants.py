"""Model registry + task types (reference utils/constants.py:32-68, inference_demo.py:54-63)."""
import importlib

BENCHMARK_REPORT_FILENAME = "benchmark_report.json"
TEST_PROMPT = "I believe the meaning of life is"

_P = "neuronx_distributed_inference_b200.models"
# model-type -> task -> "module:Class" (imported lazily)
MODEL_TYPES = {
    "llama": {"causal-lm": f"{_P}.llama.modeling_llama:NeuronLlamaForCausalLM"},
    "mistral": {"causal-lm": f"{_P}.mistral.modeling_mistral:NeuronMistralForCausalLM"},
    "qwen2": {"causal-lm": f"{_P}.qwen2.modeling_qwen2:NeuronQwen2ForCausalLM"},
    "qwen3": {"causal-lm": f"{_P}.qwen3.modeling_qwen3:NeuronQwen3ForCausalLM"},
    "gemma3": {"causal-lm": f"{_P}.gemma3.modeling_gemma3:NeuronGemma3ForCausalLM",
               "image-text-to-text": "neuronx_distributed_inference_b200.contrib.models.gemma3_vision:NeuronGemma3ForConditionalGeneration"},
    "mixtral": {"causal-lm": f"{_P}.mixtral.modeling_mixtral:NeuronMixtralForCausalLM"},
    "dbrx": {"causal-lm": f"{_P}.dbrx.modeling_dbrx:NeuronDbrxForCausalLM"},
    "qwen3_moe": {"causal-lm": f"{_P}.qwen3_moe.modeling_qwen3_moe:NeuronQwen3MoeForCausalLM"},
    "llama4": {"causal-lm": f"{_P}.llama4.modeling_llama4_text:NeuronLlama4TextForCausalLM",
               "image-text-to-text": f"{_P}.llama4.modeling_llama4:NeuronLlama4ForCausalLM"},
    "gpt_oss": {"causal-lm": f"{_P}.gpt_oss.modeling_gpt_oss:NeuronGptOssForCausalLM"},
    "deepseek": {"causal-lm": f"{_P}.deepseek.modeling_deepseek:NeuronDeepseekForCausalLM"},
    "mllama": {"image-text-to-text": f"{_P}.mllama.modeling_mllama:NeuronMllamaForCausalLM"},
    "pixtral": {"image-text-to-text": f"{_P}.pixtral.modeling_pixtral:NeuronPixtralForCausalLM"},
    "qwen2_vl": {"image-text-to-text": f"{_P}.qwen2_vl.modeling_qwen2_vl:NeuronQwen2VLForCausalLM"},
    "qwen3_vl": {"image-text-to-text": f"{_P}.qwen3_vl.modeling_qwen3_vl:NeuronQwen3VLForCausalLM"},
    "whisper": {"speech-to-text": f"{_P}.whisper.modeling_whisper:NeuronApplicationWhisper"},
    "flux": {"text-to-image": f"{_P}.diffusers.flux.application:NeuronFluxApplication"},
}
_C = "neuronx_distributed_inference_b200.contrib.models.llama_family"
MODEL_TYPES.update({
    "phi3": {"causal-lm": f"{_C}:NeuronPhi3ForCausalLM"}, "granite": {"causal-lm": f"{_C}:NeuronGraniteForCausalLM"},
    "smollm3": {"causal-lm": f"{_C}:NeuronSmolLM3ForCausalLM"}, "seed_oss": {"causal-lm": f"{_C}:NeuronSeedOssForCausalLM"},
    "olmo2": {"causal-lm": f"{_C}:NeuronOlmo2ForCausalLM"}, "olmo3": {"causal-lm": f"{_C}:NeuronOlmo3ForCausalLM"},
    "gemma2": {"causal-lm": f"{_C}:NeuronGemma2ForCausalLM"}, "glm4": {"causal-lm": f"{_C}:NeuronGlm4ForCausalLM"},
    "helium": {"causal-lm": f"{_C}:NeuronHeliumForCausalLM"}, "ernie4_5": {"causal-lm": f"{_C}:NeuronErnie4_5ForCausalLM"},
    "arcee": {"causal-lm": f"{_C}:NeuronArceeForCausalLM"}, "hunyuan_v1_dense": {"causal-lm": "neuronx_distributed_inference_b200.contrib.models.recent_families:NeuronHunYuanDenseForCausalLM"},
})
_K = "neuronx_distributed_inference_b200.contrib.models.classic_family"
MODEL_TYPES.update({
    "starcoder2": {"causal-lm": f"{_K}:NeuronStarcoder2ForCausalLM"}, "stablelm": {"causal-lm": f"{_K}:NeuronStableLmForCausalLM"},
    "cohere": {"causal-lm": f"{_K}:NeuronCohereForCausalLM"}, "gpt_neox": {"causal-lm": f"{_K}:NeuronGPTNeoXForCausalLM"},
    "gpt2": {"causal-lm": f"{_K}:NeuronGPT2ForCausalLM"}, "opt": {"causal-lm": f"{_K}:NeuronOPTForCausalLM"},
    "gptj": {"causal-lm": f"{_K}:NeuronGPTJForCausalLM"}, "phi": {"causal-lm": f"{_K}:NeuronPhiForCausalLM"},
    "falcon": {"causal-lm": f"{_K}:NeuronFalconForCausalLM"}, "gpt_bigcode": {"causal-lm": f"{_K}:NeuronGPTBigCodeForCausalLM"},
    "gpt_neo": {"causal-lm": f"{_K}:NeuronGPTNeoForCausalLM"}, "biogpt": {"causal-lm": f"{_K}:NeuronBioGptForCausalLM"},
})
_M = "neuronx_distributed_inference_b200.contrib.models.moe_family"
MODEL_TYPES.update({
    "qwen2_moe": {"causal-lm": f"{_M}:NeuronQwen2MoeForCausalLM"}, "olmoe": {"causal-lm": f"{_M}:NeuronOlmoeForCausalLM"},
    "exaone4": {"causal-lm": f"{_M}:NeuronExaone4ForCausalLM"}, "granitemoe": {"causal-lm": f"{_M}:NeuronGraniteMoeForCausalLM"},
    "phimoe": {"causal-lm": f"{_M}:NeuronPhimoeForCausalLM"}, "glm4_moe": {"causal-lm": f"{_M}:NeuronGlm4MoeForCausalLM"},
    "deepseek_v2": {"causal-lm": f"{_M}:NeuronDeepseekV2ForCausalLM"},
    "dots1": {"causal-lm": f"{_M}:NeuronDots1ForCausalLM"}, "ernie4_5_moe": {"causal-lm": f"{_M}:NeuronErnie4_5MoeForCausalLM"},
    "llava": {"image-text-to-text": "neuronx_distributed_inference_b200.contrib.models.llava:NeuronLlavaForCausalLM"},
    "qwen2_5_vl": {"image-text-to-text": "neuronx_distributed_inference_b200.contrib.models.qwen2_5_vl:NeuronQwen25VLForCausalLM"},
})
_X = "neuronx_distributed_inference_b200.contrib.models.more_families"
MODEL_TYPES.update({
    "gemma": {"causal-lm": f"{_X}:NeuronGemmaForCausalLM"}, "vaultgemma": {"causal-lm": f"{_X}:NeuronVaultGemmaForCausalLM"},
    "glm": {"causal-lm": f"{_X}:NeuronGlmForCausalLM"}, "cohere2": {"causal-lm": f"{_X}:NeuronCohere2ForCausalLM"},
    "apertus": {"causal-lm": f"{_X}:NeuronApertusForCausalLM"}, "nemotron": {"causal-lm": f"{_X}:NeuronNemotronForCausalLM"},
    "persimmon": {"causal-lm": f"{_X}:NeuronPersimmonForCausalLM"}, "xglm": {"causal-lm": f"{_X}:NeuronXGLMForCausalLM"},
    "codegen": {"causal-lm": f"{_X}:NeuronCodeGenForCausalLM"}, "openai-gpt": {"causal-lm": f"{_X}:NeuronOpenAIGPTForCausalLM"},
})
_R = "neuronx_distributed_inference_b200.contrib.models.recent_families"
MODEL_TYPES.update({"ministral": {"causal-lm": f"{_R}:NeuronMinistralForCausalLM"}, "cwm": {"causal-lm": f"{_R}:NeuronCwmForCausalLM"},
                    "olmo": {"causal-lm": f"{_R}:NeuronOlmoForCausalLM"}, "hunyuan_v1_moe": {"causal-lm": f"{_R}:NeuronHunYuanMoEForCausalLM"},
                    "flex_olmo": {"causal-lm": f"{_R}:NeuronFlexOlmoForCausalLM"}, "minimax_m2": {"causal-lm": f"{_R}:NeuronMiniMaxM2ForCausalLM"},
                    "solar_open": {"causal-lm": f"{_R}:NeuronSolarOpenForCausalLM"}, "exaone_moe": {"causal-lm": f"{_R}:NeuronExaoneMoeForCausalLM"}, "jais2": {"causal-lm": f"{_R}:NeuronJais2ForCausalLM"}, "glm4_moe_lite": {"causal-lm": f"{_R}:NeuronGlm4MoeLiteForCausalLM"},
                    "youtu": {"causal-lm": f"{_R}:NeuronYoutuForCausalLM"}, "ministral3": {"causal-lm": f"{_R}:NeuronMinistral3ForCausalLM"}, "nanochat": {"causal-lm": f"{_R}:NeuronNanoChatForCausalLM"},
                    "granitemoeshared": {"causal-lm": f"{_R}:NeuronGraniteMoeSharedForCausalLM"}})
# Hugging Face model_type spellings of families registered above under the reference's names
for _alias, _name in {"deepseek_v3": "deepseek", "gemma3_text": "gemma3", "llama4_text": "llama4", "code_llama": "llama"}.items():
    MODEL_TYPES[_alias] = {"causal-lm": MODEL_TYPES[_name]["causal-lm"]}
MODEL_TYPES["qwen3_next"] = {"causal-lm": "neuronx_distributed_inference_b200.contrib.models.qwen3_next:NeuronQwen3NextForCausalLM"}
for _t in ("qwen3_5_text", "qwen3_5_moe_text"):
    MODEL_TYPES[_t] = {"causal-lm": "neuronx_distributed_inference_b200.contrib.models.qwen3_next:NeuronQwen3_5ForCausalLM"}
_A = "neuronx_distributed_inference_b200.contrib.models.alibi_family"
MODEL_TYPES.update({"bloom": {"causal-lm": f"{_A}:NeuronBloomForCausalLM"}, "mpt": {"causal-lm": f"{_A}:NeuronMptForCausalLM"}})
_H = "neuronx_distributed_inference_b200.contrib.models.hybrid_family"
MODEL_TYPES.update({"jamba": {"causal-lm": f"{_H}:NeuronJambaForCausalLM"}, "mamba": {"causal-lm": f"{_H}:NeuronMambaForCausalLM"}, "falcon_mamba": {"causal-lm": f"{_H}:NeuronFalconMambaForCausalLM"},
                    "nemotron_h": {"causal-lm": f"{_H}:NeuronNemotronHForCausalLM"}, "lfm2_moe": {"causal-lm": f"{_H}:NeuronLfm2MoeForCausalLM"}, "bamba": {"causal-lm": f"{_H}:NeuronBambaForCausalLM"}, "mamba2": {"causal-lm": f"{_H}:NeuronMamba2ForCausalLM"}, "granitemoehybrid": {"causal-lm": f"{_H}:NeuronGraniteHybridForCausalLM"}, "lfm2": {"causal-lm": f"{_H}:NeuronLfm2ForCausalLM"}, "falcon_h1": {"causal-lm": f"{_H}:NeuronFalconH1ForCausalLM"},
                    "recurrent_gemma": {"causal-lm": f"{_H}:NeuronRecurrentGemmaForCausalLM"}})
MODEL_TYPES.update({
    "idefics": {"image-text-to-text": "neuronx_distributed_inference_b200.contrib.models.idefics:NeuronIdeficsForCausalLM"},
    "mistral3": {"image-text-to-text": "neuronx_distributed_inference_b200.contrib.models.mistral3:NeuronMistral3ForCausalLM"},
    "afmoe": {"causal-lm": f"{_M}:NeuronTrinityForCausalLM"}, "trinity": {"causal-lm": f"{_M}:NeuronTrinityForCausalLM"},
    # Chandra OCR is a Qwen3-VL fine-tune (reference contrib/models/chandra uses the stock Qwen3-VL application)
    "chandra": {"image-text-to-text": f"{_P}.qwen3_vl.modeling_qwen3_vl:NeuronQwen3VLForCausalLM"},
})
_B = "neuronx_distributed_inference_b200.contrib.models.backbone_ports"
MODEL_TYPES.update({"minicpm": {"causal-lm": f"{_B}:NeuronMiniCPMForCausalLM"}, "internlm3": {"causal-lm": f"{_B}:NeuronInternLM3ForCausalLM"},
                    "orion": {"causal-lm": f"{_B}:NeuronOrionForCausalLM"}, "janus": {"causal-lm": f"{_B}:NeuronJanusForCausalLM"},
                    "ovis2_5": {"causal-lm": f"{_B}:NeuronOvis2_5ForCausalLM"},
                    "qwen2_5_omni": {"causal-lm": f"{_B}:NeuronQwen2_5OmniForCausalLM",
                                     "audio-text-to-text": "neuronx_distributed_inference_b200.contrib.models.qwen2_5_omni:NeuronQwen2_5OmniThinkerForCausalLM"}})
MODEL_TYPES.update({"wav2vec2": {"audio-frame-classification":
                                 "neuronx_distributed_inference_b200.contrib.models.wav2vec2:NeuronWav2Vec2ForAudioFrameClassification"}})
TASK_TYPES = ("causal-lm", "image-text-to-text", "speech-to-text", "text-to-image", "audio-frame-classification", "audio-text-to-text")


def get_model_cls(model_type: str, task_type: str = "causal-lm"):
    try:
        ref = MODEL_TYPES[model_type][task_type]
    except KeyError:
        raise ValueError(f"unsupported model-type/task-type: {model_type}/{task_type}; "
                         f"known: {sorted(MODEL_TYPES)}") from None
    mod, cls = ref.split(":")
    return getattr(importlib.import_module(mod), cls)
