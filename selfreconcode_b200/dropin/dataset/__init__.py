from .dataset import SceneDataset, RandomSampler, ClipSampler, ShardedSampler, getDatasetAndLoader, write_sequence
