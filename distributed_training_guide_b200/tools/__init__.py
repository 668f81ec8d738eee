"""Command-line helpers: checkpoint consolidation, pretrained-weight loading."""
