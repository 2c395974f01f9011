DATASETS = ['sent140', 'femnist', 'shakespeare', 'celeba', 'synthetic', 'reddit', 'cifar10', 'mnist']
SEED_FILES = {'sampling': 'sampling_seed.txt', 'split': 'split_seed.txt'}
