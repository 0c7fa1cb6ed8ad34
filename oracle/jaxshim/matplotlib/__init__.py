"""Import stub so that `import iactrace` succeeds without matplotlib (display code is never run)."""
