"""ORACLE support (test infrastructure): makes the UNMODIFIED reference importable in the build
container so that the restatements in oracle/*.py can be pinned against it and golden vectors
can be generated (oracle/make_golden.py).  /root/reference does not exist on the GPU box, so
nothing under tests/ -m gpu, smoke() or bench.py imports this module.

Shims (each one is needed by the stock code):
  1. matplotlib is absent            -> stub modules before `data.utils` is imported
  2. run.py parses argv at import    -> sys.argv patched
  3. weights are CWD-relative files  -> temp work dir with seeded state_dicts saved by torch.save
  4. transformers 5.x removed the `init_weights()` entry the reference calls -> guarded alias
  5. HF hub id resolved offline      -> local directory named like the hub id inside the work dir
  6. ffmpeg/torchaudio decode absent -> convert_mp4_to_mp3 replaced by a function returning the wav
  7. 7-class driver never sets flag_save_prob (reference bug) -> class attribute
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile
import types

REFERENCE_SRC = "/root/reference/src"
HUB_ID = "audeering/wav2vec2-large-robust-12-ft-emotion-msp-dim"


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


def _stub_matplotlib() -> None:
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    colors = types.ModuleType("matplotlib.colors")
    plt.cm = types.SimpleNamespace(Blues=None)
    colors.LinearSegmentedColormap = type("LinearSegmentedColormap", (), {})
    mpl.pyplot = plt
    mpl.colors = colors
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.colors": colors})


def _patch_transformers() -> None:
    from transformers import PreTrainedModel

    if getattr(PreTrainedModel, "_avcer_oracle_patched", False):
        return
    orig = getattr(PreTrainedModel, "init_weights", None)

    def init_weights(self):
        if not hasattr(self, "all_tied_weights_keys"):
            return self.post_init()
        if orig is not None:
            return orig(self)

    PreTrainedModel.init_weights = init_weights
    PreTrainedModel._avcer_oracle_patched = True


def w2v_config(num_hidden_layers: int = 12):
    from transformers import Wav2Vec2Config

    return Wav2Vec2Config(hidden_size=1024, num_hidden_layers=num_hidden_layers, num_attention_heads=16,
                          intermediate_size=4096, conv_dim=[512] * 7, conv_kernel=[10, 3, 3, 3, 3, 2, 2],
                          conv_stride=[5, 2, 2, 2, 2, 2, 2], conv_bias=True, feat_extract_norm="layer",
                          do_stable_layer_norm=True, num_conv_pos_embeddings=128,
                          num_conv_pos_embedding_groups=16, layer_norm_eps=1e-5)


@contextlib.contextmanager
def reference_env(workdir: str | None = None):
    """Context in which `import run`, `import get_prob_video` ... resolve to the stock reference."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present (GPU box?) -- the harness only runs in the build container")
    _stub_matplotlib()
    _patch_transformers()
    old_cwd, old_argv, old_path = os.getcwd(), list(sys.argv), list(sys.path)
    tmp = None
    if workdir is None:
        tmp = tempfile.TemporaryDirectory(prefix="avcer_oracle_")
        workdir = tmp.name
    os.makedirs(os.path.join(workdir, "src", "weights"), exist_ok=True)
    os.chdir(workdir)
    sys.argv = ["run.py"]
    sys.path.insert(0, REFERENCE_SRC)
    try:
        yield workdir
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
        sys.path[:] = old_path
        if tmp is not None:
            tmp.cleanup()


def save_video_weights(workdir: str, sd_vs, sd_vd) -> None:
    import torch

    torch.save(dict(sd_vs), os.path.join(workdir, "src", "weights", "FER_static_ResNet50_AffectNet.pt"))
    torch.save(dict(sd_vd), os.path.join(workdir, "src", "weights", "FER_dinamic_LSTM_Aff-Wild2.pt"))


def reference_resnet(sd_vs):
    from architectures.video import ResNet50

    m = ResNet50(7, channels=3)
    m.load_state_dict(sd_vs, strict=True)
    return m.eval()


def reference_lstm(sd_vd):
    from architectures.video import LSTMPyTorch

    m = LSTMPyTorch()
    m.load_state_dict(sd_vd, strict=True)
    return m.eval()


def reference_audio_model(sd_a, num_classes: int = 8, num_hidden_layers: int = 12):
    if "gru.weight_ih_l0" in sd_a:
        from architectures.audio_8_cl import ExprModelV1 as cls          # the GRU variant (audio_8_cl.py:18-72)
    elif num_classes == 8:
        from architectures.audio_8_cl import ExprModelV3 as cls
    else:
        from architectures.audio_7_cl import ExprModelV2 as cls
    m = cls(w2v_config(num_hidden_layers))
    missing, unexpected = m.load_state_dict(sd_a, strict=False)
    assert not unexpected, unexpected
    assert all("masked_spec_embed" in k or k.endswith("num_batches_tracked") for k in missing), missing
    return m.eval()
