"""Training entry point with the reference's flags (main.py:9-22 of the reference):

    python main.py -c config.yaml -d <data_dir|synthetic> -train_set train_128 \
        -train_index_file train_samples_128.json -store_model_path <path> -t <tag> -iters N

Under ``torchrun --nproc-per-node N`` every rank trains on its own batches with one NCCL
gradient all-reduce per step (data parallel)."""
import os
from argparse import ArgumentParser

import torch

from adaptive_voice_conversion_b200.config import load_config
from adaptive_voice_conversion_b200.solver import Solver

if __name__ == "__main__":
    p = ArgumentParser()
    p.add_argument("-config", "-c", default="config.yaml")
    p.add_argument("-data_dir", "-d", default="synthetic")
    p.add_argument("-train_set", default="train")
    p.add_argument("-train_index_file", default="train_samples_64.json")
    p.add_argument("-logdir", default="log/")
    p.add_argument("--load_model", action="store_true")
    p.add_argument("--load_opt", action="store_true")
    p.add_argument("-store_model_path", default="model")
    p.add_argument("-load_model_path", default="model")
    p.add_argument("-summary_steps", default=100, type=int)
    p.add_argument("-save_steps", default=5000, type=int)
    p.add_argument("-tag", "-t", default="init")
    p.add_argument("-iters", default=0, type=int)
    args = p.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    solver = Solver(config=load_config(args.config), args=args)
    if args.iters > 0:
        solver.train(n_iterations=args.iters)
