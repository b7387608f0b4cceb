"""Output directory map with the reference's layout (``code/file_paths.py:7-17``).

The reference roots everything at ``<repo>/scBONITA_output`` next to its ``code/`` directory.
Here the root is ``$SCBONITA_OUTPUT`` if set, else ``./scBONITA_output`` under the current
working directory; call :func:`set_output_root` to change it at run time (the dict object is
updated in place, so modules that imported ``file_paths`` see the change)."""
import os

file_paths: dict = {}


def set_output_root(root: str, input_root: str | None = None) -> dict:
    root = os.path.abspath(root)
    input_root = os.path.abspath(input_root or os.path.join(os.path.dirname(root), "input"))
    file_paths.update({
        "pickle_files": os.path.join(root, "pickle_files"),
        "graphml_files": os.path.join(root, "graphml_files"),
        "rules_output": os.path.join(root, "rules_output"),
        "importance_score_output": os.path.join(root, "importance_score_output"),
        "relative_abundance_output": os.path.join(root, "relative_abundance_output"),
        "pathway_xml_files": os.path.join(root, "pathway_xml_files"),
        "custom_graphml": os.path.join(input_root, "custom_graphml_files"),
        "trajectories": os.path.join(root, "trajectories"),
        "trajectory_analysis": os.path.join(root, "trajectory_analysis"),
    })
    return file_paths


set_output_root(os.environ.get("SCBONITA_OUTPUT", os.path.join(os.getcwd(), "scBONITA_output")))
