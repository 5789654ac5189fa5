"""B200-native implementation of the `--version fast` demultiplexing hot path of
paulranum11/SPLiT-Seq_demultiplexing.  See DESIGN.md."""
from ._lib import (EMIT_MATCHED, EMIT_PASSING, FILTER_DISTINCT_UMI, FILTER_READS, LIB_PATH, build, exported_symbols, load,
                   needs_build)
from .distributed import demux_batches, global_filter, shard_record_ranges
from .engine import (DEFAULT_POSITIONS, Demultiplexer, HostPipeline, MalformedFastq, synth_device, synth_host, synth_sizes, synth_spec)

__all__ = ["demux_batches", "global_filter", "shard_record_ranges", "HostPipeline", "EMIT_MATCHED", "EMIT_PASSING", "FILTER_DISTINCT_UMI", "FILTER_READS", "LIB_PATH", "build", "exported_symbols",
           "load", "needs_build", "DEFAULT_POSITIONS", "Demultiplexer", "MalformedFastq", "synth_device", "synth_host",
           "synth_sizes", "synth_spec"]
