def update_dict_with_default(default_dict, update_dict):
    """reference MFGP_ver2023May/utils/dict_tools.py:3-9 (shallow merge that returns/mutates its arguments)."""
    if update_dict is None:
        return default_dict
    for key in default_dict.keys():
        if key not in update_dict.keys():
            update_dict[key] = default_dict[key]
    return update_dict
