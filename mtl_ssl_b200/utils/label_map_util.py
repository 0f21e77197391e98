"""Label map (`StringIntLabelMap` text proto) helpers: the functions of
/root/reference/object_detection/utils/label_map_util.py:25-155 the input readers, record writers and the evaluator
call (`label_map_path` of the pipeline config's input readers), on the in-tree text-format parser."""
import logging

from ..protos import text_format


def _validate_label_map(label_map):
    for item in label_map.item:
        if item.id < 1:
            raise ValueError("Label map ids should be >= 1.")


def create_category_index(categories):
    """{id: category dict}."""
    return {cat["id"]: cat for cat in categories}


def convert_label_map_to_categories(label_map, max_num_classes, use_display_name=True):
    """-> [{'id', 'name'}] for ids in 1..max_num_classes, first item wins per id; no label map -> 'category_<id>'."""
    categories = []
    if not label_map:
        return [{"id": i + 1, "name": "category_{}".format(i + 1)} for i in range(max_num_classes)]
    seen = []
    for item in label_map.item:
        if not 0 < item.id <= max_num_classes:
            logging.info("Ignore item %d since it falls outside of requested label range.", item.id)
            continue
        name = item.display_name if use_display_name and item.HasField("display_name") else item.name
        if item.id not in seen:
            seen.append(item.id)
            categories.append({"id": item.id, "name": name})
    return categories


def load_labelmap(path_or_text):
    """Text-format file (or the text itself) -> StringIntLabelMap message."""
    text = str(path_or_text)
    if "{" not in text:                              # a path
        with open(text) as f:
            text = f.read()
    label_map = text_format.Merge(text, text_format.Message("StringIntLabelMap"))
    _validate_label_map(label_map)
    return label_map


def get_label_map_dict(label_map_path):
    """{name: id}."""
    return {item.name: item.id for item in load_labelmap(label_map_path).item}


def get_class_indices(label_map_dict):
    return sorted(label_map_dict.values())


def get_index_map_dict(label_map_dict):
    out = {v: k for k, v in label_map_dict.items()}
    out[0] = "bg"
    return out


def categories_from_input_reader(input_reader, max_num_classes):
    """evaluator / eval.py: `label_map_path` of an input reader -> categories list for the metric functions."""
    return convert_label_map_to_categories(load_labelmap(input_reader.label_map_path), max_num_classes)
