"""Train / evaluate / sample driver — same entry point and flags as the reference
(`python -um train.train --data ... --model ... --task ... --checkpt_dir ... --init_dir ...`,
reference src/train/train.py:36-42, README.md:58-64), same config merge order data -> task ->
model (:49-51), same model registry (:12-15), same loop structure (:79-126).

Differences, all defaulting to reference behaviour: `yaml.safe_load` (PyYAML 6), argument parsing
inside main() instead of at import, `episodes_per_step` episodes per optimizer step, and — when
launched under torchrun — one process per GPU with the episodes of a step sharded across ranks
(one NCCL all-reduce of the flat gradient buffer per step inside the engine).
"""
import argparse
import os
import pprint
from importlib import import_module

import yaml

from data.episode import load_sampler_from_config

PP = pprint.PrettyPrinter(depth=6)


def load_model_from_config(config):
    Model = getattr(import_module(config['model_module_name']), config['model_class_name'])
    return Model(config)


def write_seq(seq, dir, name):
    if isinstance(seq, str):
        with open(os.path.join(dir, name + '.txt'), 'w') as text_file:
            text_file.write(seq)
    else:
        seq.write(os.path.join(dir, name + '.mid'))


def evaluate(model, episode_sampler, n_episodes, rank=0, world=1, shared_stream=True):
    """Mean of per-episode mean NLLs (reference :27-33).

    Under torchrun the n_episodes of ONE evaluation are sharded over the ranks and the sum is all-reduced, so an N-GPU run
    reports the number the reference's loop would.  With `shared_stream` (val / test samplers: same seed on every rank) every
    rank draws all n_episodes from the identical sampler stream and scores episodes i = rank (mod world): exactly the
    episodes of the single-process evaluation.  Without it (the train sampler, seeded per rank) each rank scores its share
    of episodes from its own stream."""
    if world <= 1:
        avg_nll = 0.
        for _ in range(n_episodes):
            avg_nll += model.eval(episode_sampler.get_episode())
        return avg_nll / n_episodes
    total = 0.
    for i in range(n_episodes):
        if shared_stream:
            episode = episode_sampler.get_episode()        # keeps the stream aligned across ranks
            if i % world == rank:
                total += model.eval(episode)
        elif i % world == rank:
            total += model.eval(episode_sampler.get_episode())
    return all_reduce_sum(total) / n_episodes


def all_reduce_sum(value):
    import torch
    import torch.distributed as dist
    device = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return float(t[0])


def build_parser():
    parser = argparse.ArgumentParser(description='Train a model.')
    for flag in ('data', 'model', 'task', 'checkpt_dir', 'init_dir'):
        parser.add_argument('--' + flag, dest=flag, default='')
    return parser


def load_config(args):
    config = yaml.safe_load(open(args.data, 'r'))
    config.update(yaml.safe_load(open(args.task, 'r')))
    config.update(yaml.safe_load(open(args.model, 'r')))
    config['dataset_path'] = os.path.abspath(config['dataset_path'])
    config['checkpt_dir'] = args.checkpt_dir
    return config


def init_distributed():
    """torchrun launch: one process per GPU, NCCL over NVLink."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group('nccl')
    return dist.get_rank(), world


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank, world = init_distributed()
    log = print if rank == 0 else (lambda *a, **k: None)
    log('Args:')
    log(PP.pformat(vars(args)))
    config = load_config(args)
    log('Config:')
    log(PP.pformat(config))

    episode_sampler = {}
    base_seed = config.get('seed', None)
    for split in config['splits']:
        config['split'] = split
        if base_seed is not None and world > 1 and split == 'train':   # every rank trains on its own episodes;
            config['seed'] = base_seed + rank                            # val / test keep ONE stream (sharded in evaluate)
        elif base_seed is not None:
            config['seed'] = base_seed
        episode_sampler[split] = load_sampler_from_config(config)
    if base_seed is not None:
        config['seed'] = base_seed  # identical initial weights on every rank

    config['input_size'] = episode_sampler['train'].get_num_unique_words()
    if not config['input_size'] > 0:
        raise RuntimeError('error reading data: %d unique tokens processed' % config['input_size'])
    log('Num unique words: %d' % config['input_size'])

    n_train, print_every_n, val_every_n = config['n_train'], config['print_every_n'], config['val_every_n']
    n_val, n_test, n_samples, max_len = config['n_val'], config['n_test'], config['n_samples'], config['max_len']
    per_step = int(config.get('episodes_per_step', 1))

    model = load_model_from_config(config)
    model.recover_or_init(args.init_dir)

    shared = base_seed is not None      # unseeded samplers (np.random) cannot be aligned across ranks
    avg_nll = evaluate(model, episode_sampler['val'], n_val, rank, world, shared)
    log('Iter: %d, val-nll: %.3e' % (0, avg_nll))

    avg_loss = 0.
    for i in range(1, n_train + 1):
        if per_step == 1:
            loss = model.train(episode_sampler['train'].get_episode())
        else:
            loss = model.train([episode_sampler['train'].get_episode() for _ in range(per_step)])
        avg_loss += loss
        if i % val_every_n == 0:
            avg_nll = evaluate(model, episode_sampler['val'], n_val, rank, world, shared)
            log('Iter: %d, val-nll: %.3e' % (i, avg_nll))
            if args.checkpt_dir != '':
                model.save(args.checkpt_dir)
        if i % print_every_n == 0:
            log('Iter: %d, loss: %.3e' % (i, avg_loss / print_every_n))
            avg_loss = 0.

    for split, label in (('train', 'Train'), ('val', 'Validation'), ('test', 'Test')):
        log('%s Avg NLL: %.3e' % (label, evaluate(model, episode_sampler[split], n_test, rank, world, shared and split != 'train')))

    if rank == 0:
        write_samples(model, episode_sampler, args, n_samples, max_len)
    if world > 1:       # nobody leaves (and tears the process group down) while rank 0 is still sampling
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def write_samples(model, episode_sampler, args, n_samples, max_len):
    samples_dir = os.path.join(args.checkpt_dir, 'samples')
    os.makedirs(samples_dir, exist_ok=True)
    for i in range(n_samples):
        curr_sample_dir = os.path.join(samples_dir, 'sample_%d' % i)
        os.makedirs(curr_sample_dir, exist_ok=True)
        episode = episode_sampler['test'].get_episode()
        support_set = episode.support[0]
        sample = model.sample(support_set, max_len)
        for j in range(support_set.shape[0]):
            write_seq(episode_sampler['test'].detokenize(support_set[j]), curr_sample_dir, 'support_%d' % j)
        write_seq(episode_sampler['test'].detokenize(sample), curr_sample_dir, 'model_sample')


if __name__ == '__main__':
    main()
