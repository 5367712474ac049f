"""cacophony_b200 — B200-native (sm_100a) implementation of Cacophony's inference hot path.

Mirrors the export list of the reference's ``src/caco_torch/__init__.py`` plus the frontend functions of
``src/eval/eval_caco_torch.py``; all arithmetic runs in ``libcaco_b200.so`` (C ABI: ``include/caco_b200.h``).
"""
from .model import (CACO, CACOConfig, AudioAttentionPooler, create_caco_model, AudioEncoder, AudioTransformerConfig,
                    RobertaModel, RobertaConfig, RobertaDecoder, NORM_EPS)
from .checkpoint import convert_caco_checkpoint
from .frontend import DatasetConfig, compute_mel_spectrogram, spectrogram_to_patches, prepare_audio_batch
from .loader import pad_ragged, prepare_audio_batch_ragged, resample_to_16k, load_audio
from .eval import (load_caco_torch, prepare_text_batch, compute_audio_embedding, compute_text_embedding,
                   compute_all_class_embeddings, zs_classification, audio_retrieval, compute_retrieval_metric,
                   decode_caption, decode_caption_ids)

__all__ = [
    "CACO", "CACOConfig", "AudioAttentionPooler", "create_caco_model", "convert_caco_checkpoint",
    "AudioEncoder", "AudioTransformerConfig",
    "RobertaModel", "RobertaConfig", "RobertaDecoder", "NORM_EPS", "DatasetConfig", "compute_mel_spectrogram", "spectrogram_to_patches",
    "prepare_audio_batch", "pad_ragged", "prepare_audio_batch_ragged", "resample_to_16k", "load_audio", "load_caco_torch",
    "prepare_text_batch", "compute_audio_embedding", "compute_text_embedding", "compute_all_class_embeddings",
    "zs_classification", "audio_retrieval", "compute_retrieval_metric", "decode_caption", "decode_caption_ids",
]
