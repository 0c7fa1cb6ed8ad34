"""Import stub so that `import iactrace` succeeds without trimesh (display code is never run)."""
