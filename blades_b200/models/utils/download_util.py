"""Google-Drive download helper (reference models/utils/download_util.py).  Needs network access."""
import requests

_URL = "https://docs.google.com/uc?export=download"


def get_confirm_token(response):
    return next((v for k, v in response.cookies.items() if k.startswith('download_warning')), None)


def save_response_content(response, destination, chunk_size=32768):
    with open(destination, "wb") as f:
        for chunk in response.iter_content(chunk_size):
            if chunk:
                f.write(chunk)


def download_file_from_google_drive(id, destination):
    session = requests.Session()
    response = session.get(_URL, params={'id': id}, stream=True)
    token = get_confirm_token(response)
    if token:
        response = session.get(_URL, params={'id': id, 'confirm': token}, stream=True)
    save_response_content(response, destination)
