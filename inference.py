"""One-shot voice conversion entry point with the reference's flags (inference.py:95-109).
The wav<->mel DSP is not part of this build; pass mels as .npy ([T, n_mels], already
normalised when -attr is omitted) to convert without a vocoder:

    python inference.py -c config.yaml -m model.ckpt -s src.npy -t tgt.npy -o out.npy
"""
from argparse import ArgumentParser

import numpy as np
import torch

from adaptive_voice_conversion_b200.config import load_config
from adaptive_voice_conversion_b200.inference import Inferencer
from adaptive_voice_conversion_b200.utils import local_device

if __name__ == "__main__":
    p = ArgumentParser()
    p.add_argument("-attr", "-a", help="attr file path")
    p.add_argument("-config", "-c", help="config file path")
    p.add_argument("-model", "-m", help="model path")
    p.add_argument("-source", "-s", help="source mel .npy")
    p.add_argument("-target", "-t", help="target mel .npy")
    p.add_argument("-output", "-o", help="output mel .npy")
    p.add_argument("-sample_rate", "-sr", default=24000, type=int)
    args = p.parse_args()
    inf = Inferencer(config=load_config(args.config), args=args)
    src, tgt = np.load(args.source).astype(np.float32), np.load(args.target).astype(np.float32)
    if inf.attr is not None:
        src, tgt = inf.normalize(src), inf.normalize(tgt)
    dev = local_device()
    _, mel = inf.inference_one_utterance(torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev))
    np.save(args.output, mel)
