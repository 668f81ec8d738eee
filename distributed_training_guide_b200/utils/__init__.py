"""CLI, logging, timers, memory stats, data pipeline and checkpointing shared by all chapters."""
